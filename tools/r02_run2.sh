set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
timeout 300 python tools/e2e_diag.py > gpurun_out/r02b_e2e_diag.txt 2>&1
VOLSTN_HOST_RAMP=0 timeout 300 python tools/e2e_diag.py > gpurun_out/r02b_e2e_diag_noramp.txt 2>&1
timeout 300 python tools/kbench.py --grids g1,g2,g3 --algos 0,6 --iters 20 > gpurun_out/r02b_kbench_tma.txt 2>&1
timeout 300 python tools/kbench.py --grids g2,g5 --algos 0,6,2 --iters 10 --size 128 --batch 8 >> gpurun_out/r02b_kbench_tma.txt 2>&1
timeout 300 python tools/fbench.py --what cage --iters 20 > gpurun_out/r02b_fbench_cage.txt 2>&1
timeout 300 python tools/fbench.py --what cage --iters 20 --csz 4 >> gpurun_out/r02b_fbench_cage.txt 2>&1
ls -la gpurun_out | tail
