import sys, os, tempfile, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import conftest, redtime_b200 as rt
with tempfile.TemporaryDirectory() as tmp:
    d1 = conftest.make_example1_dir(os.path.join(tmp,'a'))
    d2 = conftest.make_example1_dir(os.path.join(tmp,'b'), switches=[1,0,1,1])
    out={}
    for tag,d,cfg in (('nk256_1loop',d1,dict(nk=256)),('nk256_full',d2,dict(nk=256)),('hiacc_full',d2,dict(nk=256,beta_kmin=1e-5,beta_kmax=20.0,n_lnk=1000,a_early=1e-50))):
        h=rt.RedTimeB200(**cfg); h.add_cosmology(rt.read_run_dir(d)); h.prepare(); t,_,_,st=h.run(); out[tag]=t[0]; print(tag,st,h.counters(0)); h.close()
    np.savez_compressed('gpurun_out/nk256_tables.npz',**out)
