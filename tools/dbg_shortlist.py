"""Tuning aid: run one shortlist query with GLOC_DEBUG_SHORTLIST=1 and print list statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GLOC_DEBUG_SHORTLIST"] = "1"
import numpy as np
import gloc3d_b200 as g
from gloc3d_b200 import synth
n, nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 10000
db = synth.make_descriptors(n, seed=1234, dup_run=8)
q = np.concatenate([synth.make_queries(db, nq // 2, seed=5678), synth.make_queries(db, nq - nq // 2, seed=5679, sigma=0.01)])
ix = g.KnnIndex(512, 0); ix.set_db(db); ix.set_mode(g.KNN_SHORTLIST)
for _ in range(2):
    t = time.time(); idx, d2 = ix.query(q, 25); print("query s", time.time() - t)
st = ix.stats(); print("rows/query", st.shortlist_rows / max(1, st.shortlist_queries), "fallback", st.fallback_queries)
