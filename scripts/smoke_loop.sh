#!/bin/bash
# Repeats the driver's smoke in fresh processes (round-1 driver smoke failed once at HEAD; hunting a flake).
N=${1:-20}
for i in $(seq 1 $N); do
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|Error|error" | tr '\n' ' '
  echo
done
