import importlib.util,sys
spec=importlib.util.spec_from_file_location('s','profiles/summarize.py'); sys.argv=['x','zz']; m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
for rep in sys.argv[1:] if False else []: pass
import sys as _s
