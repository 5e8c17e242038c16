#!/bin/bash
# Runs a gtest-shim binary built from the reference's test/testTensor.cu the way the reference's CI does:
# from <root>/build/test, with <root>/python/b_d.bt present (ref: ci/script.sh:24-46, testTensor.cu:194).
# usage: scripts/run_gtest_binary.sh <binary> [gtest args...]
set -euo pipefail
BIN=$(realpath "$1"); shift
# optional wrapper form: run_gtest_binary.sh <tool> <tool args...> <binary>: resolve relative paths before the cd
ARGS=(); for a in "$@"; do if [ -e "$a" ]; then ARGS+=("$(realpath "$a")"); else ARGS+=("$a"); fi; done; set -- "${ARGS[@]+"${ARGS[@]}"}"
REPO=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d)
mkdir -p "$WORK/python" "$WORK/build/test"
( cd "$REPO" && python -c "
import sys; sys.path.insert(0, 'oracle')
import bt_format
bt_format.write_bt('$WORK/python/b_d.bt', bt_format.reference_b_d())" )
cd "$WORK/build/test"
set +e
"$BIN" "$@"
RC=$?
set -e
rm -rf "$WORK"
exit $RC
