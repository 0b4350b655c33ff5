#!/bin/bash
# Round-2 ncu captures, one gpurun call (every profiled command first runs once without ncu):
#   gpurun --timeout 2400 -- 'bash tools/ncu_round2.sh'
# Outputs land in gpurun_out/ and are summarised here (CPU) with tools/ncu_summary.py / tools/launch_shares.py.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() {   # run <name> <kernel regex> <skip> <count> <command...>
  local name=$1 k=$2 s=$3 c=$4; shift 4
  "$@" > $O/r02_${name}_plain.log 2>&1 && $NCU -k regex:$k -s $s -c $c -o $O/r02_${name} -f "$@" > $O/r02_${name}_ncu.log 2>&1
  echo "$name: plain rc=$? $(tail -n 1 $O/r02_${name}_plain.log)"
}
# 1. the fused step's launch list (one steady-state step between cudaProfilerStart/Stop)
python bench.py --warmup 2 --profile-one-step > $O/r02_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_bench_step.csv \
    python bench.py --warmup 2 --profile-one-step > $O/r02_step_ncu.log 2>&1
echo "launch list rc=$?"
# 2. the dominant kernel: fused 1x1 convolution + BN + residual + ReLU, (256, 64 -> 256, 56 x 56)
run conv_64_256_res conv1x1_tc_kernel 3 1 python tools/profile_conv.py 64 256 56 1
run conv_512_128 conv1x1_tc_kernel 3 1 python tools/profile_conv.py 512 128 28 0
# 3. Gram solver on ResNet-50 layer 14 (512 x 256 x 200960): tensor-core Gram kernel and the recurrence kernel
run gram_tc_512x256x200960 gram_tc_kernel 1 1 python tools/profile_layer.py 512 256 200960 2 1
run gram_path_512x256x200960 gram_path_kernel 1 1 python tools/profile_layer.py 512 256 200960 2 1
# 4. direct solver, multi-launch structure (256 x 1024 x 12800): sweep + recurrence of one 32-feature block
run sweep_256x1024x12800 sweep_kernel 40 1 env GPFQ_RESIDENT=0 python tools/profile_layer.py 256 1024 12800 2 0
run recur_256x1024x12800 recur_kernel 40 1 env GPFQ_RESIDENT=0 python tools/profile_layer.py 256 1024 12800 2 0
# 5. direct solver, resident structure (512 x 4608 x 768)
run resident_512x4608x768 resident_kernel 1 1 python tools/profile_layer.py 512 4608 768 2 0
ls -la $O/*.ncu-rep | tail -12
