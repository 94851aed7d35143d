#!/bin/bash
# Runs the reference's whole Python test directory twice -- on the unmodified reference library and on the drop-in
# librebound linked against the MOCK engine (tests/hostlogic, CPU, backed by the oracle) -- and prints both summaries.
# Test infrastructure for the authoring container (needs /root/reference); see DESIGN.md section 4.
# usage: tools/reference_tests_on_mock.sh [resident-mode: "" | 0 | 1]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
make -C "$ROOT/tests/hostlogic" > /dev/null
W=$(mktemp -d)
mkdir -p "$W/ref" "$W/mock"
cp "$ROOT/oracle/_ref/libref_harness.so" "$W/ref/librebound.so"
cp "$ROOT/tests/hostlogic/_build/librebound.so" "$ROOT/tests/hostlogic/_build/librebound_b200.so" "$W/mock/"
IGN="--ignore=$REF/rebound/tests/test_horizons.py --ignore=$REF/rebound/tests/test_plotting.py --ignore=$REF/rebound/tests/test_server.py"
cd "$W"
echo "== reference library"
OMP_NUM_THREADS=1 PYTHONPATH="$W/ref:$REF" python -m pytest "$REF/rebound/tests" -q -p no:cacheprovider --no-header $IGN 2>&1 | tail -6 || true
echo "== drop-in on the mock engine (REBOUND_B200_RESIDENT='$1')"
REBOUND_B200_RESIDENT="$1" OMP_NUM_THREADS=1 LD_LIBRARY_PATH="$W/mock:$ROOT/oracle" PYTHONPATH="$W/mock:$REF" \
    python -m pytest "$REF/rebound/tests" -q -p no:cacheprovider --no-header $IGN 2>&1 | tail -6 || true
rm -rf "$W"
