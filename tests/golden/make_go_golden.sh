#!/bin/bash
# Produce the Go-side golden vectors with the UNMODIFIED reference.  Needs a Go toolchain (absent from the build
# image of this repository, so the outputs are not committed yet).  Usage:
#   tests/golden/make_go_golden.sh /path/to/rtlamr-checkout
# Writes tests/golden/go_dump_*.jsonl (decoder level) and go_msgs_*.jsonl (message level); tests/test_go_golden.py
# then pins the oracle (and, on a GPU box, the CUDA path) against them.
set -euo pipefail
REF=${1:?rtlamr checkout}
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
WORK=$(mktemp -d)
cp -r "$REF" "$WORK/rtlamr"
cd "$WORK/rtlamr"
cp "$REPO/go/goldengen/protocol/golden_dump_test.go" protocol/
mkdir -p cmd/goldengen && cp "$REPO/go/goldengen/cmd/goldengen/main.go" cmd/goldengen/
if ! go list golang.org/x/xerrors >/dev/null 2>&1; then   # offline: stub the only third-party import of the path
  mkdir -p stubs && cp -r "$REPO/go/goldengen/stubs/xerrors" stubs/
  go mod edit -replace golang.org/x/xerrors=./stubs/xerrors
fi
# synthetic streams of this repository's generator (same bytes on every machine): written by the python side first
python3 "$HERE/make_golden.py" --synthetic-inputs "$WORK"
run() {  # name input msgtypes chiplength
  GOLDEN_IN=$2 GOLDEN_OUT="$HERE/go_dump_$1.jsonl" GOLDEN_MSGTYPES=$3 GOLDEN_CL=$4 go test ./protocol -run TestDumpGolden -count=1
  go run ./cmd/goldengen -in "$2" -msgtype "$3" -symbollength "$4" > "$HERE/go_msgs_$1.jsonl"
}
# absolute paths: `go test ./protocol` runs the test binary inside the package directory
run sample_cl78_scm "$PWD/assets/sample.bin" scm 78
run sample_cl72_scm "$PWD/assets/sample.bin" scm 72
run synth_cl72_scm "$WORK/synth_cl72_scm.bin" scm 72
run synth_cl72_multi "$WORK/synth_cl72_multi.bin" scm,scm+,idm 72
run synth_cl72_r900 "$WORK/synth_cl72_r900.bin" r900 72
run synth_cl32_all "$WORK/synth_cl32_all.bin" scm,scm+,idm,r900 32
echo "wrote $(ls "$HERE"/go_*.jsonl | wc -l) files under $HERE"
