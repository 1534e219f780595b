import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lm_ab
from fiksi_b200 import workloads as wl
for n in (256, 512, 1024, 2048, 3072):
    lm_ab.run("truss20", wl.truss(n), reps=20)
    lm_ab.run("cad_mix", wl.cad_mix(n), reps=20)
    lm_ab.run("hinged4", wl.hinged_triangles(4, n), reps=20)
