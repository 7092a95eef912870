#!/bin/bash
# round 2, GPU call 3: kernel v2 (register double buffer, 16-byte cells, one-warp fast path): parity, slot sweep for two
# unroll factors, ncu launch list + full capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_acct.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02c_pytest_acct.log
python profiles/scripts/r02_sweep.py U8 > gpurun_out/r02c_sweep_U8.jsonl 2> gpurun_out/r02c_sweep_U8.err
make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc VARIANT=-DGEMHOOK_UNROLL=16 > gpurun_out/r02c_build16.log 2>&1
python profiles/scripts/r02_sweep.py U16 > gpurun_out/r02c_sweep_U16.jsonl 2> gpurun_out/r02c_sweep_U16.err
make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc VARIANT=-DGEMHOOK_UNROLL=4 > gpurun_out/r02c_build4.log 2>&1
python profiles/scripts/r02_sweep.py U4 > gpurun_out/r02c_sweep_U4.jsonl 2> gpurun_out/r02c_sweep_U4.err
make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc > /dev/null 2>&1
tail -4 gpurun_out/r02c_pytest_acct.log
for t in U8 U16 U4; do echo "== $t"; python - <<PY
import json
for l in open("gpurun_out/r02c_sweep_$t.jsonl"):
    d=json.loads(l); print(d["nslots"], d["n"], d["env"], d["ms"], d["gbps"], d["frac"], d["grid"])
PY
tail -2 gpurun_out/r02c_sweep_$t.err; done
