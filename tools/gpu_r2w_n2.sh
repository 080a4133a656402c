#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "rows_pack" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/dbg_sparse.py 2>&1 | grep "rank " | head -20
