set -u
mkdir -p gpurun_out
for n in 1024 4096; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/small_$n.csv python scripts/small_batch_profile.py $n 2>&1 | tail -1
done
for c in edge128 edge256; do
MORIG_LIB=$PWD/morig_b200/libmorig_b200_trace.so python scripts/tc_trace.py $c f16 5 > gpurun_out/trace_$c.txt 2>&1
tail -3 gpurun_out/trace_$c.txt | cut -c1-1500
done
