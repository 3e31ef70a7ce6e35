import time, torch, sys
sys.path.insert(0,'/root/repo')
from wekws_b200 import init_model, model_config, synth
for name,B,T in (('gru',512,1),('mdtc',1,40),('mdtc',1024,40)):
    m = synth.randomize_(init_model(model_config(name))).eval().cuda()
    x = synth.features(B,T,80).cuda()
    c = torch.zeros(2,B,128,device='cuda') if name=='gru' else torch.zeros(B,64,244,device='cuda')
    for _ in range(20): y,c = m(x,c)
    torch.cuda.synchronize()
    N=500
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(N): y,c = m(x,c)
    e1.record(); t_enq=time.perf_counter()-t0
    torch.cuda.synchronize(); t_all=time.perf_counter()-t0
    print(name,B,T,'enqueue us/call',round(t_enq/N*1e6,1),'gpu us/call (events over N)',round(e0.elapsed_time(e1)/N*1e3,1),'wall',round(t_all/N*1e6,1))
