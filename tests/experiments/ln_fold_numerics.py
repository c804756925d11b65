import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import oracle
from oracle import encoder as E
from oracle.signals import speech_like
from oracle import logmel

def rnd(x): return x.to(torch.bfloat16).float()

def folded_linear(x, ln_w, ln_b, W, b):
    # x: bf16-valued fp32 [T,d]; emulate: stats fp32, GEMM on raw x with W' = bf16(gamma*W), fp32 accumulate
    mean = x.mean(-1, keepdim=True)
    var = ((x-mean)**2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var+1e-5)
    Wp = rnd(W*ln_w[None,:])
    colsum = Wp.sum(-1)
    bp = b + W @ ln_b
    acc = x @ Wp.T
    return rstd*(acc - mean*colsum[None,:]) + bp[None,:]

orig_forward = E.encoder_forward
def run(cfgname, T, fold):
    cfg = oracle.CONFIGS[cfgname]
    w = oracle.make_weights(cfg, seed=1)
    w = {k: torch.as_tensor(v) for k,v in w.items()}
    clips=[speech_like(t*160, 50+i) for i,t in enumerate(T)]
    mels=[torch.from_numpy(logmel(c)).to(torch.bfloat16).float().numpy() for c in clips]
    ref,_ = E.encoder_forward(w,cfg,mels,emulate_bf16=False)
    if not fold:
        out,_ = E.encoder_forward(w,cfg,mels,emulate_bf16=True)
    else:
        # monkeypatch: replace F.layer_norm+_linear pairs: implement by patching _linear and layer_norm through a flag
        state={}
        real_ln=F.layer_norm; real_lin=E._linear
        def fake_ln(x, shape, weight, bias, eps):
            state['x']=x; state['w']=weight; state['b']=bias
            class Tag(torch.Tensor): pass
            y=real_ln(x,shape,weight,bias,eps)
            state['y_id']=None
            return y
        def fake_lin(x, weight, bias, fp8):
            if 'x' in state and x.shape==state['x'].shape and state.get('pending',False):
                pass
            return real_lin(x,weight,bias,fp8)
        # simpler: re-implement the layer loop here
        d=cfg.d_model
        x, toks, inter = E.encoder_forward(w,cfg,mels,emulate_bf16=True,return_intermediate=True)
        xx = inter['embed']
        wins=[]
        for m,n in zip(mels,toks): wins += E.window_lens(n,cfg,m.shape[1])
        for li in range(cfg.layers):
            p=f"layers.{li}."
            # attention with folded LN1 for q,k,v
            lnw,lnb=w[p+"self_attn_layer_norm.weight"],w[p+"self_attn_layer_norm.bias"]
            q=rnd(folded_linear(xx,lnw,lnb,w[p+"self_attn.q_proj.weight"],w[p+"self_attn.q_proj.bias"]))
            k=rnd(folded_linear(xx,lnw,lnb,w[p+"self_attn.k_proj.weight"],w[p+"self_attn.k_proj.bias"]))
            v=rnd(folded_linear(xx,lnw,lnb,w[p+"self_attn.v_proj.weight"],w[p+"self_attn.v_proj.bias"]))
            H=cfg.heads; hd=d//H
            outs=[]; s=0
            for wl in wins:
                qq=q[s:s+wl].view(wl,H,hd).transpose(0,1); kk=k[s:s+wl].view(wl,H,hd).transpose(0,1); vv=v[s:s+wl].view(wl,H,hd).transpose(0,1)
                att=torch.softmax(qq@kk.transpose(1,2)*hd**-0.5,-1)
                o=(rnd(att)@vv).transpose(0,1).reshape(wl,d)
                outs.append(o); s+=wl
            o=rnd(torch.cat(outs,0))
            o=rnd(F.linear(o,w[p+"self_attn.out_proj.weight"],w[p+"self_attn.out_proj.bias"]))
            xx=rnd(xx+o)
            lnw,lnb=w[p+"final_layer_norm.weight"],w[p+"final_layer_norm.bias"]
            h=rnd(folded_linear(xx,lnw,lnb,w[p+"fc1.weight"],w[p+"fc1.bias"]))
            h=rnd(F.gelu(h))
            h=rnd(F.linear(h,w[p+"fc2.weight"],w[p+"fc2.bias"]))
            xx=rnd(xx+h)
        h=rnd(folded_linear(xx,w["ln_post.weight"],w["ln_post.bias"],w["proj1.weight"],w["proj1.bias"]))
        h=rnd(F.gelu(h))
        out=rnd(F.linear(h,w["proj2.weight"],w["proj2.bias"]))
    r=ref.numpy(); o=out.numpy()
    return np.abs(o-r).max()/(r.max()-r.min()), np.abs(o-r).max()/np.abs(r).max()
for name,T in (("tiny",[300,177,1234]),("0.6B",[500])):
    print(name,'std  ',run(name,T,False))
    print(name,'fold ',run(name,T,True))
