// Package xerrors: offline stand-in for golang.org/x/xerrors, enough for csv/csv.go:30 (xerrors.Errorf).
// Use with `replace golang.org/x/xerrors => ./stubs/xerrors` in go.mod when the module cache is empty.
package xerrors

import "fmt"

func Errorf(format string, a ...interface{}) error { return fmt.Errorf(format, a...) }
