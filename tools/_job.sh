timeout 600 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12
tools/ab_bench.sh godotoceanwaves_b200/libocean_prev.so godotoceanwaves_b200/libocean.so
tools/ncu_capture.sh r02_b3 godotoceanwaves_b200/libocean.so
