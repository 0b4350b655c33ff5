#!/bin/bash
# Round-2 (second session) ncu captures after the operand swap of the convolution kernel; every profiled command first
# runs once without ncu:   gpurun --timeout 1800 -- 'bash tools/ncu_round2b.sh'
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
run() {   # run <name> <kernel regex> <skip> <count> <command...>
  local name=$1 k=$2 s=$3 c=$4; shift 4
  "$@" > $O/r02b_${name}_plain.log 2>&1 && $NCU -k regex:$k -s $s -c $c -o $O/r02b_${name} -f "$@" > $O/r02b_${name}_ncu.log 2>&1
  echo "$name: plain rc=$? $(tail -n 1 $O/r02b_${name}_plain.log)"
}
python bench.py --warmup 2 --profile-one-step > $O/r02b_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02b_launches_bench_step.csv \
    python bench.py --warmup 2 --profile-one-step > $O/r02b_step_ncu.log 2>&1
echo "launch list rc=$?"
run conv_64_256_res conv1x1_tc_kernel 3 1 python tools/profile_conv.py 64 256 56 1
run conv_256_64 conv1x1_tc_kernel 3 1 python tools/profile_conv.py 256 64 56 0
run conv_1024_256 conv1x1_tc_kernel 3 1 python tools/profile_conv.py 1024 256 14 0
ls -la $O/r02b_*.ncu-rep
