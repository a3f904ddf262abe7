"""Two-config run-to-run determinism check of the production path (GPU box)."""
import os, sys
sys.argv = [sys.argv[0]]
import runpy
os.environ.setdefault('REPS', '6')
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'determinism_diag.py')).read()
head = src[:src.index("ref = run(")]
exec(compile(head, 'determinism_diag.py', 'exec'))
run('default', {})
run('pairs off, stream-K off, graphs off', {'LUMI_CONV_2CTA': '0', 'LUMI_CONV_STREAMK': '0', 'LUMI_GRAPHS': '0'})
run('default, pipeline off', {}, pipeline=False)
