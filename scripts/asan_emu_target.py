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
# one bucket owning more than 64 threads of a round: the block-filled start-bucket list (aff_start_bucket_heavy_kernel)
eng.set_msm_affine(2); eng.set_msm_window(0)
big=np.tile(b[20:120],(26,1)); sv=F.fr_to_mont_array([3]*big.shape[0])
assert (eng.msm(sv,big)[0]==c.msm_pippenger(sv,big)).all()
eng.set_msm_affine(-1)
# batched evaluations (different lengths, one empty) and the division-step batch inverse
vs=[eng.vec(F.random_fr_mont(40+k,m)) for k,m in enumerate((1,33,9000))]+[eng.vec(n=0)]
xs=[3,5,7,11]
got=eng.fr_eval(vs,xs)
from oracle import pyref
for v,x,g in zip(vs,xs,got):
    assert g==pyref.poly_eval(v.ints(),x)
inv=eng.vec(F.random_fr_mont(50,3000)); want=[pow(a,-1,pyref.R) for a in inv.ints()]
eng.fr_batch_inverse(inv); assert inv.ints()==want
d=F.random_fr_mont(8,1<<12); e=d.copy(); eng.ntt(e,12); assert (e==c.ntt(d,12)).all()
d=F.random_fr_mont(9,1<<13); e=d.copy(); eng.ntt(e,13); assert (e==c.ntt(d,13)).all(); eng.ntt(e,13,inverse=True); assert (e==d).all()
eng.close(); print("asan run OK")
