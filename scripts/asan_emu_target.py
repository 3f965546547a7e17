import sys, numpy as np
sys.path.insert(0,'/root/repo')
import zkp_implementation_b200 as z
from oracle import coracle as c
F=z.fields
eng=z.Engine(0, lib_path=sys.argv[1])
n=777
b=c.srs(F.fr_to_mont_array([0xAFF1]), n); b[11]=0
s=F.random_fr_mont(91,n); s[7]=0
for rounds in (1,3,6):
    eng.set_msm_affine(rounds)
    for w in (4,9,0):
        eng.set_msm_window(w)
        out,inf=eng.msm(s,b)
        assert (out==c.msm_pippenger(s,b)).all()
    eng.set_msm_window(0)
    eng.srs_upload(b); eng.srs_precompute(6)
    assert (eng.msm(s)[0]==c.msm_pippenger(s,b)).all()
    sv=F.fr_to_mont_array([5]*n)
    assert (eng.msm(sv)[0]==c.msm_pippenger(sv,b)).all()
eng.set_msm_affine(-1)
d=F.random_fr_mont(8,1<<12); e=d.copy(); eng.ntt(e,12); assert (e==c.ntt(d,12)).all()
d=F.random_fr_mont(9,1<<13); e=d.copy(); eng.ntt(e,13); assert (e==c.ntt(d,13)).all(); eng.ntt(e,13,inverse=True); assert (e==d).all()
eng.close(); print("asan run OK")
