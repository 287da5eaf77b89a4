import sys; sys.path.insert(0,'/root/repo')
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200.dmd import DMD
tab = tables.load_default_tables(); topo, sv = genconfig.system_b(tab, 0.18, seed=1)
for eng in (1,2):
    d = DMD(tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=4, engine=eng), topo, tab); d.set_state(sv)
    d.run(20000); s0=d.stats(0); d.run(50000); s1=d.stats(0)
    ev=s1.events-s0.events
    print("engine",eng,"pp/ev",(s1.pair_predictions-s0.pair_predictions)/ev,"nv/ev",(s1.nbr_visits-s0.nbr_visits)/ev)
