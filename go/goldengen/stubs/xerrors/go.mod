module golang.org/x/xerrors

go 1.21
