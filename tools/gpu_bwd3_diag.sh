#!/bin/bash
for cp in 416 424 400; do echo "== colp $cp"; SWINB200_BWD3_COLP=$cp timeout 300 python tools/bwd3_diag.py 2>&1 | tail -5; done
