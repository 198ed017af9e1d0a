set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
GOLEMFLAVOR_B200_LIB=scratch/variants/lib_head.so python scratch/k2_bench.py gpurun_out/k2_ref.npy
python scratch/k2_bench.py gpurun_out/k2_ref.npy
GOLEMFLAVOR_B200_LIB=scratch/variants/lib_head.so python scratch/scan_bench.py 1e9 texture,anarchic
python scratch/scan_bench.py 1e9 texture,anarchic
