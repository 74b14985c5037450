#!/bin/bash
mkdir -p gpurun_out
echo "== probe small"; timeout -s KILL 120 python scripts/umma_probe.py 24 30 24 3 2>&1 | tail -6; echo "rc=$?"
echo "== probe big"; timeout -s KILL 120 python scripts/umma_probe.py 120 214 120 6 2>&1 | tail -6; echo "rc=$?"
