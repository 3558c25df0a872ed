import sys, json
sys.path.insert(0,'/root/repo/tests/golden')
from make_golden import import_reference
import_reference()
from rlcard.utils.utils import reorganize
import random
rng=random.Random(3)
cases=[]
for _ in range(20):
    P=rng.randint(1,4)
    traj=[]
    for p in range(P):
        k=rng.randint(0,5)
        t=[]
        for i in range(k):
            t += ['s%d_%d'%(p,i), rng.randint(0,9)]
        t.append('s%d_T'%p)
        traj.append(t)
    pay=[rng.choice([-1,0,1,0.5,2]) for _ in range(P)]
    cases.append({'trajectories':traj,'payoffs':pay,'expected':reorganize(traj,pay)})
json.dump(cases, open('/root/repo/tests/golden/reorganize.json','w'))
print(len(cases))
