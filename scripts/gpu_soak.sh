# Soak: many random shapes / options against the oracle chain (tests/test_gpu_fuzz.py), several seeds.
for seed in 1 2 3; do PWGS_FUZZ_CASES=${1:-150} PWGS_FUZZ_SEED=$seed timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4; done
