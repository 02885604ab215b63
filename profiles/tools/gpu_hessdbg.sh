#!/bin/bash
for d in 0 8 1 2 4 6 7 3 5; do echo "dbg=$d"; MPX_HESS_DBG=$d timeout 300 python profiles/tools/evaluators_time.py 2>/dev/null | grep "hess_l" | cut -c1-120; done
