import numpy as np, sys
sys.path.insert(0,'/root/repo')
from optuna_b200 import TPEEngine, ParamSpec
from oracle import motpe as mo
from tests._util import load
g=load('motpe.npz')
eng=TPEEngine(0)
v=g['mo3/v']; n=v.shape[0]
rs=np.random.RandomState(0)
eng.set_space([ParamSpec(kind=0,low=0.0,high=1.0) for _ in range(2)])
eng.set_history(rs.uniform(0,1,(n,2)), np.zeros(n,np.int8), np.zeros((n,2)))
eng.set_values(v,0)
for nb in (10,25,33,40,50,60,64):
    info=eng.prepare([0,1], n_below=nb, n_candidates=8, multivariate=True)
    below,_=eng.get_split()
    want=mo.split_complete_mo(v,nb)
    eng.build()
    w=eng.get_mo_weights(); ref=mo.weights_below_mo(v[want])
    print(nb, np.array_equal(below,want), np.abs(w-ref).max(), int(mo.is_pareto_front(v[want],False).sum()))
