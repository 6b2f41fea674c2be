#!/bin/bash
# usage (on the GPU box, under gpurun): tools/prof_ncu.sh NAME KERNEL_REGEX <prof_one.py args...>
# plain run first, then one `ncu --set full` capture of the kernel; the report comes back xz-compressed
name=$1; regex=$2; shift 2
mkdir -p /tmp/ncu_out gpurun_out
python tools/prof_one.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -o /tmp/ncu_out/$name -f \
    python tools/prof_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
xz -T8 -c /tmp/ncu_out/$name.ncu-rep > gpurun_out/$name.ncu-rep.xz
tail -1 gpurun_out/plain_$name.log
