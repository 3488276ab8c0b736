import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metmhn_b200 import Handle
G = np.load('tests/golden/golden_v1.npz')
cases = sys.argv[1:] or ['kinds_n1', 'kinds_n3']
for c in cases:
    th, dp, dm, dat = [G[f'{c}/{k}'] for k in ('theta','d_p','d_m','dat')]
    params = np.concatenate([th.ravel(), dp, dm]); n_tot = th.shape[0]; sq = n_tot*n_tot
    for r in range(dat.shape[0]):
        h = Handle(dat[r:r+1]); s, g = h.eval_weighted(params, 1.0, 1.0); h.close()
        ref = G[f'{c}/row_logp'][r]
        eg = np.abs(g[:sq].reshape(n_tot,n_tot) - G[f'{c}/row_g'][r]).max()
        ep = np.abs(g[sq:sq+n_tot] - G[f'{c}/row_gdp'][r]).max()
        em = np.abs(g[sq+n_tot:] - G[f'{c}/row_gdm'][r]).max()
        print(c, r, dat[r].tolist(), 'lp', s, ref, 'err', abs(s-ref), 'g', eg, 'dp', ep, 'dm', em)
        if max(eg, ep, em) > 1e-9 and '-v' in os.environ.get('DBG',''):
            print(g[:sq].reshape(n_tot,n_tot)); print(G[f'{c}/row_g'][r]); print(g[sq:], G[f'{c}/row_gdp'][r], G[f'{c}/row_gdm'][r])
