import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racer_tracer_b200 import harness as H, capi
from oracle import oracle as O
cfg = H.load_config(os.path.join(ROOT, 'tests/golden/config.yml'))
r = H.CudaRenderer([0])
for name in ['three_balls','emissive','noise_and_textures','cornell_box','clown']:
    for prec in (64, 32):
        w=h=600
        job = H.prepare_job(os.path.join(ROOT, f'tests/golden/scenes/{name}.yml'), cfg, w, h)
        r.upload(job)
        p = H.make_params(w,h,1,20,fixed_jitter=1)
        ids,t,nrm,pt = r.primary_aov(p, prec)
        oids,ot,onrm,opt = O.primary_aov(job,p)
        bad = np.nonzero(ids != oids)[0]
        hit = (oids != 0) & (ids == oids)
        rel = np.abs(t[hit]-ot[hit])/np.abs(ot[hit]) if hit.any() else np.zeros(1)
        print(name, prec, 'id diffs', len(bad), 't rel max %.3e' % rel.max(), 'n err %.3e' % np.abs(nrm[hit]-onrm[hit]).max())
        for i in bad[:12]:
            print('   px', i % w, i // w, 'gpu id', ids[i], 't', t[i], 'oracle id', oids[i], 't', ot[i], 'opt', opt[i], 'gpt', pt[i])
