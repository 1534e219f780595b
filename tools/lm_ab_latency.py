import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lm_ab
from fiksi_b200 import workloads as wl
for n in (1, 32, 64):
    lm_ab.run("hinged16", wl.hinged_triangles(16, n), reps=20)
    lm_ab.run("truss20", wl.truss(n), reps=20)
    lm_ab.run("truss40", wl.truss(n, n_points=40), reps=20)
    lm_ab.run("cad_mix", wl.cad_mix(n), reps=20)
