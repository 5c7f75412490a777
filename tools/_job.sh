timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
tools/ab_bench.sh godotoceanwaves_b200/libocean_prev.so godotoceanwaves_b200/libocean.so godotoceanwaves_b200/libocean_keepb.so godotoceanwaves_b200/libocean.so godotoceanwaves_b200/libocean_keepb.so
