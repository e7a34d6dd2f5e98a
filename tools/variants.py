"""Build A/B variants of libamtfeat.so with extra -D flags:  python tools/variants.py name=DEF1,DEF2 name2=...
Outputs amt_tools_b200/variants/<name>.so; run with AMTFEAT_LIB=<path>."""
import importlib.util
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('_b', os.path.join(ROOT, 'amt_tools_b200', 'build.py'))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
os.makedirs(os.path.join(ROOT, 'amt_tools_b200', 'variants'), exist_ok=True)


def one(arg):
    name, _, defs = arg.partition('=')
    out = os.path.join(ROOT, 'amt_tools_b200', 'variants', name + '.so')
    b.build(force=True, out=out, defines=[d for d in defs.split(',') if d])
    return out


with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
