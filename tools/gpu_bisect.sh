#!/bin/bash
for v in "" astg; do echo "== '$v'"; MMT_B200_DEV_LIB=$v timeout 600 python tools/attn_check_shapes.py 2>&1 | tail -4; done
