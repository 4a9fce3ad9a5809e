#!/bin/bash
# First GPU bring-up: every step in its own process with its own timeout, so one faulting
# kernel cannot poison the rest.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name: $*"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "    exit $?"; tail -n 12 gpurun_out/$name.log | sed 's/^/    /'; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
run ops python -m pytest tests/test_gpu_parity.py -q -k "op_level or bound_the_way or errors_are" --timeout 250
run peak python -m pytest tests/test_gpu_parity.py -q -s -k "tensor_pipe_peak" --timeout 120
for spec in "8 3" "6 3" "5 3" "7 2" "1 2" "2 3" "3 9" "4 33"; do
  set -- $spec
  run probe_tf32_L$1 python tools/probe_block.py $1 $2 tf32
done
for spec in "8 3" "6 3" "1 2" "2 3" "3 9"; do
  set -- $spec
  run probe_bf16_L$1 python tools/probe_block.py $1 $2 bf16
done
run probe_L0 python tools/probe_block.py 0 2 tf32
run fp32_e2e python -m pytest tests/test_gpu_parity.py -q -s -k "golden and fp32" --timeout 250
run debug_mode python -m pytest tests/test_gpu_parity.py -q -s -k "debug_mode" --timeout 250
