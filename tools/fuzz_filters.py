#!/usr/bin/env python
"""Random filters (node counts 2..400, even / random / clustered sampling, zeroed wings, noisy responses) through the host
emulation of the device algorithm against the oracle, both modes: `python tools/fuzz_filters.py [seed] [trials]`
(CPU only; FUZZ_GPU=1 runs the CUDA path instead; prints every entry above tolerance and the worst errors)."""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import cases, emul
from egg_b200 import synth
from oracle.oracle import Oracle
np.seterr(all='ignore')
L=cases.libs(); opt, ir = L['opt'], L['ce01']
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
worst={'fp32c':0,'fp64':0}
underflow=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 20):
    bands=[]
    for b in range(6):
        n=int(rng.integers(2,400))
        c=10**rng.uniform(-0.4,2.3)           # centre 0.4 .. 200 um
        wid=c*10**rng.uniform(-2.5,-0.3)
        kind=rng.integers(0,4)
        if kind==0: lam=np.linspace(c-wid/2,c+wid/2,n)
        elif kind==1: lam=np.sort(rng.uniform(c-wid/2,c+wid/2,n))
        elif kind==2: lam=c-wid/2+wid*np.sort(rng.random(n))**3
        else:
            lam=np.sort(np.concatenate([np.linspace(c-wid/2,c+wid/2,max(n//2,2)), rng.uniform(c-wid/2,c+wid/2,n-max(n//2,2))]))
        lam=np.unique(lam)
        if len(lam)<2: continue
        x=(lam-c)/(wid/2)
        res=np.clip(1-x**2,0,None)**rng.uniform(0.2,3)*(1+0.3*rng.standard_normal(len(lam)))
        if rng.random()<0.3: res[:int(rng.integers(0,len(lam)//2+1))]=0
        if rng.random()<0.3: res=np.abs(res)
        if not np.any(res!=0): res[len(res)//2]=1.0
        bands.append((lam,res))
    gal=synth.draw_galaxies(40,opt_lib=opt,ir_lib=ir,mmin=7.5,seed=int(rng.integers(1,1<<30)))
    opts={}
    if rng.random()<0.3: opts['igm']='none'
    if rng.random()<0.3: opts['no_nebular']=True
    if len(sys.argv)>3:  # wider option sampling
        for k in ('trim_filters','filter_normalize','filter_flambda','filter_photons','no_dust'):
            if rng.random()<0.25: opts[k]=True
        if rng.random()<0.2: opts['igm']='madau95'
        if rng.random()<0.2: opts['line_profile']='average'
        if rng.random()<0.2: opts['opt_lib_imf']='chabrier'
    try:
        rf=[bands[i] for i in rng.choice(len(bands),2,replace=False)] if len(sys.argv)>3 and rng.random()<0.5 else []
        O=Oracle(opt,ir,bands,rf,line_lambda=synth.LINE_LAMBDA,**opts)
    except Exception as e:
        print('oracle rejects',e); continue
    ref=O.compute(gal,f64=True)
    for prec in ('fp32c','fp64'):
        try:
            if os.environ.get('FUZZ_GPU'):   # the CUDA path itself instead of its host emulation
                import egg_b200
                ph=egg_b200.Photometry(opt,ir,bands,rf,line_lambda=synth.LINE_LAMBDA,precision=prec,**opts)
                o=ph.compute(gal,f64=True); ph.close()
                out={k:o[k+'_f64'] for k in ('flux_disk','flux_bulge')+(('rfmag_disk','rfmag_bulge') if rf else ())}
            else:
                out=emul.compute(opt,ir,bands,rf,gal,line_lambda=synth.LINE_LAMBDA,precision=prec,**opts)
        except Exception as e:
            print('emul rejects',opts,str(e)[:80]); continue
        for k in ('flux_disk','flux_bulge'):
            e=cases.rel_err(out[k],ref[k],cases.FLOOR[prec])
            if prec=='fp32c':
                # the fp32 mode forms S*W in float: a flux below ~1e-30 uJy (responses of 1e-40, which (1-x^2)^p produces at the
                # ends of a two-node band) underflows into float denormals; such entries are counted, not judged
                tiny=np.abs(ref[k])<1e-30
                underflow+=int(np.count_nonzero(tiny & (e>cases.TOL[prec])))
                e=np.where(tiny,0.0,e)
            if e.max()>worst[prec]: worst[prec]=e.max()
            if e.max()>cases.TOL[prec]:
                w=np.unravel_index(np.argmax(e),e.shape)
                print('FAIL',trial,prec,k,e.max(),'band',O.band_order[w[1]],'n',len(bands[O.band_order[w[1]]][0]),'z',gal['z'][w[0]])
        if rf:
            for k in ('rfmag_disk','rfmag_bulge'):
                dm=np.abs(np.asarray(out[k])-ref[k]); dm=dm[np.isfinite(dm)]
                if dm.size and dm.max()>(1e-5 if prec=='fp32c' else 1e-9): print('FAIL rf',trial,prec,k,dm.max())
print('worst',worst,'; fp32c entries below 1e-30 uJy off by more than the tolerance (float underflow, not judged):',underflow)
