#!/bin/bash
# First GPU call of the next round (run from the repo root through gpurun, one GPU):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash benchmarks/round2_first_call.sh'
# 1. everything written after round 1's GPU budget was spent: XPASS = green, then drop the xfail marker;
# 2. the verified suite, unchanged paths;
# 3. bench line + the opt-in variants that only need a measurement to become defaults;
# 4. the reference's own CUDA path on the same GPU (informational baseline).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_pending.py -m gpu -q -rxX --timeout 900 > gpurun_out/r2_pending.log 2>&1
python -m pytest tests -m gpu -x -q --deselect tests/test_zz_gpu_pending.py > gpurun_out/r2_tests.log 2>&1
python bench.py --steps 16 --warmup 4 > gpurun_out/r2_bench.log 2>&1
python bench.py --steps 8 --warmup 4 --no-cpu-baseline --e2e-pipelined 8 > gpurun_out/r2_bench_e2e_pipelined.log 2>&1
python bench.py --steps 16 --warmup 4 --no-cpu-baseline --e2e-steps 0 --p2g-sweep 5 > gpurun_out/r2_bench_sweep5.log 2>&1   # packed fp32 (FFMA2) sweep vs the default line above
python bench.py --impl reference-cuda --config C2 --steps 5 --warmup 2 > gpurun_out/r2_refcuda_c2.log 2>&1
python bench.py --impl reference-cuda --config C3 --steps 3 --warmup 1 > gpurun_out/r2_refcuda_c3.log 2>&1
python bench.py --config C2 --steps 16 --warmup 4 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_bench_c2.log 2>&1
tail -3 gpurun_out/r2_pending.log gpurun_out/r2_tests.log
cut -c1-400 gpurun_out/r2_bench_sweep5.log | tail -1
cut -c1-400 gpurun_out/r2_bench.log gpurun_out/r2_refcuda_c2.log gpurun_out/r2_refcuda_c3.log
python benchmarks/variants.py --combos 4:1,5:1,4:1,5:1 --tag r2_sweep5 > gpurun_out/r2_variants_sweep5.jsonl 2> gpurun_out/r2_variants_sweep5.err   # A/B on one resident workload, interleaved
tail -4 gpurun_out/r2_variants_sweep5.jsonl | cut -c1-300
# 5. C5 up to 2^30 keys (BASELINE "1M-1B keys"; round 1 measured up to 2^28)
python benchmarks/prims_sweep.py --min-log2 20 --max-log2 30 > gpurun_out/r2_prims_sweep.jsonl 2> gpurun_out/r2_prims_sweep.err
tail -2 gpurun_out/r2_prims_sweep.jsonl | cut -c1-300
# 6. C5 "vs reference CudaExecutionPolicy": the reference's own policy (CUB underneath, sync(true)) and the same generic calls on b200_exec()
for lg in 20 22 24 26 28; do python -m oracle.refcuda_runner prims-bench $lg 5; done > gpurun_out/r2_prims_vs_refcuda.jsonl 2> gpurun_out/r2_prims_vs_refcuda.err
tail -1 gpurun_out/r2_prims_vs_refcuda.jsonl | cut -c1-400
# 7. the weakly compressible fluid (EquationOfStateConfig, SURVEY §8(d) "cheap variant") on the binned path: 161.5 B / particle-substep
python bench.py --model eos --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_eos.log 2>&1
cut -c1-300 gpurun_out/r2_bench_eos.log | tail -1
# 8. memcheck over the kernels that ran for the first time today (small cases only; slow under the tool, hence the timeout)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_pending.py -m gpu -q -x \
  -k "plastic_model_matches or cuboid or index_buckets or g2p2g or grid_momentum_functors or vonmises_on_the_binned or equation_of_state_on or plastic_models_on" \
  > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck.log
# 9. racecheck on the shared-memory-staged binned kernels, one small case
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mpm.py -m gpu -q -x \
  -k "binned_path_matches_oracle and cube8" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_racecheck.log
