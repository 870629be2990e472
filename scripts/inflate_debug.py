import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
bwa = importlib.import_module("network-aware-bwa_b200")
from test_bgzf import inflate_cases
api = bwa.api
api.init([0])
for name, packed, want in inflate_cases():
    try:
        got, ooff, ms = api.bgzf_inflate(packed)
        print(name, "ok" if got == want else "DIFFERENT", len(got), len(want))
    except Exception as e:
        print(name, "ERR", str(e)[-80:])
api.destroy()
