"""Dev helper: build kernel variants (-D switches) into hoomd-tf_b200/lib/variants/ for A/B timing."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "hoomd-tf_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
outdir = os.path.join(ROOT, "hoomd-tf_b200", "lib", "variants"); os.makedirs(outdir, exist_ok=True)
for spec_ in sys.argv[1:]:
    name, _, defs = spec_.partition(":")
    out = os.path.join(outdir, "libhtf_%s.so" % name)
    b.build(defines=[d for d in defs.split(",") if d], out=out)
    print(out)
