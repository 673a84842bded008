#!/bin/bash
python scripts/gpu_debug_paths.py 2>&1 | tail -22
PIXIE_NO_GRAPH=1 python scripts/gpu_debug_paths.py 2>&1 | tail -22
