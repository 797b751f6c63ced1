#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (training path, post-process, point ops, graph prep, TMA narrow kernel)
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -x --tb=line \
    tests/test_gpu_training.py -k "linear_fwd or column_slices or batchnorm or edge_gather or segmax or pooling or normalize_and or modules_match or kat" \
    > gpurun_out/sanitizer_train.txt 2>&1
echo "train rc=$?"; tail -4 gpurun_out/sanitizer_train.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -x --tb=line \
    tests/test_gpu_postproc.py tests/test_gpu_deform.py -k "not 4096 and not pipeline" > gpurun_out/sanitizer_post.txt 2>&1
echo "post rc=$?"; tail -4 gpurun_out/sanitizer_post.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -x --tb=line \
    tests/test_gpu_parity.py -k "graph_prep or narrow or batch_equals or kat or edge_conv_motion_module" > gpurun_out/sanitizer_fwd.txt 2>&1
echo "fwd rc=$?"; tail -4 gpurun_out/sanitizer_fwd.txt
