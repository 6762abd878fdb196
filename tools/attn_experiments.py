"""What-if timing of the attention kernel with single phases removed (results are WRONG by construction; timing only)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = '''
import sys, torch; sys.path.insert(0, %r)
import tools.kernel_bench as kb
from tests.gpu_util import ptr, stream
from flow_factory_b200 import _lib
B,S,H=8,4429,24
qkv=torch.randn(B,S,3*64*H,device="cuda").bfloat16(); o=torch.empty(B,S,64*H,device="cuda",dtype=torch.bfloat16)
ms=kb.timeit(lambda: _lib.check(_lib.lib().ffb200_attention(ptr(qkv),B,S,H,ptr(o),stream())))
print(ms)
''' % ROOT
for v in ("NONE", "NO_MAX", "NO_EXP", "NO_OWAIT", "NO_STORE"):
    env = dict(os.environ, FFB200_LIB=os.path.join(ROOT, "flow_factory_b200", f"libffb200_exp_{v}.so"))
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    print(v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
