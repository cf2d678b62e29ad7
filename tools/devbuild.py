"""Development build of librdfwi.so with only the production instantiations of the cluster kernel (OpenFWI pitch, 13 rows):
python tools/devbuild.py [out.so] [-DNAME ...]   then   RDFWI_LIB=<out.so> python ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from red_diffeq_b200 import _cabi
out = next((a for a in sys.argv[1:] if a.endswith(".so")), os.path.join(ROOT, "build_tmp", "librdfwi_dev.so"))
defs = ["RDFWI_DEV_FAST"] + [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
os.makedirs(os.path.dirname(out), exist_ok=True)
t0 = time.time()
_cabi.build(force=True, defines=defs, out=out, verbose="-v" in sys.argv)
print(f"built {out} in {time.time() - t0:.0f} s with {defs}")
