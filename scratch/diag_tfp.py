import sys, os, importlib
sys.path[:0] = ['/root/repo', '/root/repo/oracle', '/root/repo/tests']
import numpy as np, helpers, fixtures, oracle_py
case_mod = importlib.import_module('fest3d_b200.case')
solver = importlib.import_module('fest3d_b200.solver')
blocks = fixtures.load(case_mod, 'tests/golden/tfp', scheme=dict(scheme_name="ausmUP", interpolant="muscl", time_step_accuracy="RK4"), control=dict(CFL=0.5))
s = solver.Solver(blocks); w = oracle_py.OracleWorld(blocks)
err, ro = w.residual(1); rg = s.residual()
for b, blk in enumerate(blocks):
    sc = helpers.flux_scale(w, b, blk)
    for v in range(7):
        e = np.abs(rg[b][v]-ro[b][v]); floor = max(sc[v].max(),1e-300)*1e-6
        rel = e/np.maximum(sc[v], floor)
        idx = np.unravel_index(np.argmax(rel), rel.shape)
        print('blk',b,'var',v,'maxrel %.2e'%rel.max(),'at (k,j,i)',idx,'abs err %.3e'%e[idx],'scale %.3e'%sc[v][idx],'res %.3e'%ro[b][v][idx], 'u=%.3e'%blk.qp[1][idx[0]+3,idx[1]+3,idx[2]+3])
