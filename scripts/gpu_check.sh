#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests group by group, each under its own timeout so a
# hanging kernel cannot eat the whole call.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run geometry 600 python -m pytest tests/test_gpu_geometry.py -m gpu -q -x
run kpconv_fp32 600 python -m pytest tests/test_gpu_kpconv.py -m gpu -q -k "fp32 or pools or int32"
run gemm_tc 180 python -m pytest tests/test_gpu_kpconv.py -m gpu -q -k "gemm_tc"
run kpconv_tc 600 python -m pytest tests/test_gpu_kpconv.py -m gpu -q -k "bf16 or seeded or linearity"
run lifting 600 python -m pytest tests/test_gpu_lifting.py -m gpu -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
