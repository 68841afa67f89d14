import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import madrona_learn_b200 as m
from madrona_learn_b200.engine import PolicyProgram
from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
from oracle import nn as onn, ppo as oppo, algo_common as oac
DEV='cuda:0'
buckets=[4,8,5,5,2,2]; A=6
D,H,L,rows=64,256,3,4096
for clip in (0.2, 10.0):
  for dt in (torch.float32, torch.bfloat16):
    rng=np.random.default_rng(1)
    p=onn.init_params(rng,D,H,L,buckets)
    p['actor']['kernel']=(rng.standard_normal(p['actor']['kernel'].shape)*0.2).astype(np.float32)
    ac=m.ActorCritic(backbone=m.BackboneShared(prefix=None,encoder=m.BackboneEncoder(net=m.models.MLP(H,L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(buckets)),critic=m.models.DenseLayerCritic())
    prog=PolicyProgram(ac,D,{'act':m.DiscreteActionsConfig(buckets)},DEV,dt)
    prog.load_oracle_params(p)
    Tp,M=4,rows//4
    cfg=oppo.PPOCfg(buckets,entropy_coef=0.02,clip_coef=clip)
    mb=dict(obs=rng.standard_normal((Tp,M,D)).astype(np.float32),
        actions=np.stack([rng.integers(0,b,(Tp,M)) for b in buckets],-1).astype(np.int32),
        advantages=rng.standard_normal((Tp,M,1)).astype(np.float32),
        returns=rng.standard_normal((Tp,M,1)).astype(np.float32),
        values=rng.standard_normal((Tp,M,1)).astype(np.float32),mb_weights=np.ones((M,1),np.float32))
    lg,cr,_=onn.actor_critic_fwd(onn.cast_tree(p,np.float64),mb['obs'].reshape(rows,D).astype(np.float64))
    lp0,_=onn.action_stats(lg,mb['actions'].reshape(rows,A),buckets)
    mb['log_probs']=(lp0+0.2*rng.standard_normal(lp0.shape)).reshape(Tp,M,A).astype(np.float32)
    ref=oppo.ppo_loss(p,mb,cfg,None,dtype=np.float64)
    dv={k:torch.from_numpy(v).to(DEV) for k,v in mb.items()}
    obs_d=dv['obs'].view(rows,D)
    head=prog.forward_train(obs_d,rows)
    tw=prog.train_ws(rows)
    mean,rstd=oac.zscore_stats(mb['advantages'])
    adv_mr=torch.tensor([mean,rstd,0,0],dtype=torch.float32,device=DEV)
    obj_scale=(ctypes.c_float*A)(*[1.0/(rows*A)]*A); ent_scale=(ctypes.c_float*A)(*[cfg.entropy_coef/(rows*A)]*A)
    prog.zero_grads()
    call('mlb_ppo_loss_f32',ptr(head),c_int(prog.NH),ptr(dv['actions']),ptr(dv['log_probs']),ptr(dv['advantages']),ptr(dv['returns']),ptr(None),ptr(None),ptr(adv_mr),ptr(None),prog._buckets_c,obj_scale,ent_scale,c_int(A),c_ll(rows),c_ll(M),c_float(cfg.clip_coef),c_float(cfg.value_loss_coef),c_int(prog.loss_flags),ptr(tw['dhead']),ptr(prog.head_bias_grad()),ptr(tw['stats_out']),ptr(tw['loss_ws']),c_size_t(tw['loss_ws'].numel()), None, c_int(0))
    prog.backward(obs_d,rows)
    g=prog.to_oracle_params(prog.grads)
    dh=tw['dhead'].cpu().numpy()
    print(f'clip={clip} dtype={dt}: dlogits rel', np.linalg.norm(dh[:,:26]-ref['dlogits'])/np.linalg.norm(ref['dlogits']))
    def rel(x,y): return np.linalg.norm(x-y)/max(np.linalg.norm(y),1e-12)
    for i in range(L):
        print('  layer',i,'kernel',rel(g['mlp'][i]['kernel'],ref['grads']['mlp'][i]['kernel']),'scale',rel(g['mlp'][i]['scale'],ref['grads']['mlp'][i]['scale']),'bias',rel(g['mlp'][i]['bias'],ref['grads']['mlp'][i]['bias']))
    print('  actor',rel(g['actor']['kernel'],ref['grads']['actor']['kernel']),'critic',rel(g['critic']['kernel'],ref['grads']['critic']['kernel']))
