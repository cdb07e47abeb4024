#!/bin/bash
VG_ATTN_TRACE=1 IMAGES=1024 ONLY=attention timeout 120 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vilgod_b200.engine import Engine
eng = Engine(num_views=4)
qkv = torch.randn(1024, 197, 2304, device="cuda").bfloat16(); qkv[:, :, :768] *= 0.35
eng.test_attention(qkv); torch.cuda.synchronize()
print("---- second call"); sys.stdout.flush()
eng.test_attention(qkv); torch.cuda.synchronize()
PY
