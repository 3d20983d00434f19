set -x
mkdir -p gpurun_out
T=${TAG:-r02ab}
which compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 60 python tools/quick_bin.py --batch 4 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_memcheck_bin.log 2>&1
tail -15 gpurun_out/${T}_memcheck_bin.log
timeout 900 compute-sanitizer --tool memcheck python tools/quick_evrep.py --batch 2 --steps 1 --only-tiled > gpurun_out/${T}_memcheck_evrep.log 2>&1
tail -15 gpurun_out/${T}_memcheck_evrep.log
