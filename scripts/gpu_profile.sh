#!/bin/bash
# Evidence run (round 2): (1) launch list of the timed (graph-replayed) bench steps; (2) full ncu capture (source attached)
# of one launch of each dominant kernel shape + the narrow kernel with and without TMA staging; (3) SASS digest.
set -u
mkdir -p gpurun_out
MORIG_BENCH_PROFILE=graph timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 --train-steps 0 > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_kernel|edge_mma" -o /tmp/prof_kernels python scripts/prof_kernels.py e256 e128 d840 e32 e16 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
ncu -i /tmp/prof_kernels.ncu-rep --page details > gpurun_out/prof_kernels_details.txt 2>/dev/null
MORIG_NARROW_TMA=0 timeout 600 ncu --set full --clock-control none --profile-from-start off \
    -k regex:"edge_mma" -o /tmp/prof_narrow_ldg python scripts/prof_kernels.py e32 > gpurun_out/ncu_narrow_ldg.log 2>&1
ncu -i /tmp/prof_narrow_ldg.ncu-rep --page raw --csv > gpurun_out/prof_narrow_ldg_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/prof_kernels.ncu-rep)
echo "report bytes: $sz"
if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_kernels.ncu-rep gpurun_out/; fi
cuobjdump -sass morig_b200/libmorig_b200.so > /tmp/sass.txt 2>/dev/null
for m in UTCHMMA UTCHMMA.2CTA LDTM UBLKCP UTMALDG UTCBAR HMMA.16816 FFMA2 FADD2 ACQBULK SYNCS.ARRIVE; do
  echo "$m $(grep -c "$m" /tmp/sass.txt)"; done > gpurun_out/sass_digest.txt
cat gpurun_out/sass_digest.txt
ls -la gpurun_out/ | tail -5
