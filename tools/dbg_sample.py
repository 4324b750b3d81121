import sys, importlib, torch, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import bruno_oracle as O
name='en2_noconv_mnist_tp'
spec=O.spec_for(name); P=O.init_params(spec,seed=13,dtype=torch.float64,perturb_process=True)
x=O.synthetic_batch(spec,1,3,seed=14).float().double()
O.data_dependent_init(spec,P,x); P=O.cast_params(O.cast_params(P,torch.float32),torch.float64)
cfg=importlib.import_module('bruno_b200.config_rnn.'+name)
xg=x.float().cuda(); cfg.build_model(xg[:1]); cfg.model.load_tf_variables(P)
T,n,D=3,4,784
g=torch.Generator().manual_seed(3)
noise=[torch.rand(2,n,1,D,generator=g) for _ in range(T+1)]
with torch.no_grad():
    xs_o,lp_o=O.build_model_sampling(spec,P,x,[t.double() for t in noise])
xs,lp=cfg.build_model(xg,sampling_mode=True,noise=[t.cuda() for t in noise])
print('oracle lp',lp_o); print('gpu lp',lp)
