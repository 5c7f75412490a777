#!/bin/bash
# scratch: the gpurun job of the current tuning batch (run as `gpurun -- bash tools/_job.sh`; see tools/ab_quick.py, tools/final_capture.sh)
