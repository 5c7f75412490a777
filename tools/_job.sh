timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -5
timeout 600 python tools/run_configs.py 2>&1 | grep -E "cfg5|spectrum generation" | cut -c1-330
