set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -k "tma" > gpurun_out/r02i_pytest_tma.log 2>&1; tail -3 gpurun_out/r02i_pytest_tma.log
for v in "" variants/libvolstn_b200_tma16c4.so variants/libvolstn_b200_tma20.so variants/libvolstn_b200_tma24.so; do
  echo "== lib ${v:-default (box 16, ctas 4/3)}" >> gpurun_out/r02i_kbench_tma_variants.txt
  if [ -n "$v" ]; then export VOLSTN_B200_LIB=$GRAFT_REPO_ROOT/alignet_b200/csrc/$v; else unset VOLSTN_B200_LIB; fi
  timeout 300 python tools/kbench.py --grids g1,g2,g5 --algos 6 --iters 20 >> gpurun_out/r02i_kbench_tma_variants.txt 2>&1
done
unset VOLSTN_B200_LIB
cat gpurun_out/r02i_kbench_tma_variants.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
