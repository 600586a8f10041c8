#!/bin/bash
for d in 0 1 2 3 4 8 16 24 28 31; do SWINB200_BWD3_DEBUG=$d timeout 120 python tools/bwd3_time.py 2>&1 | grep "(0, 0)"; done
