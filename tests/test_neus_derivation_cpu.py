"""The hand-derived backward of the NeuS compositing kernel (csrc/neus.cu: neus_bwd_kernel) checked against autograd on CPU in
fp64.  The loop below is a line-by-line Python mirror of the kernel's algorithm: recompute the forward quantities of every
section, walk the sections in reverse with the suffix sum S_i = sum_{k>i} dL/dw_k w_k,
dL/dalpha_i = dL/dw_i T_i - S_i / (1 - alpha_i + 1e-7), then the chain through clip / the two sigmoids / iter_cos / the eikonal
term, with and without the background mix.  (The kernel itself is tested on the GPU against the real reference renderer:
tests/test_neus_gpu.py.)"""
import torch


def test_neus_composite_backward_derivation_matches_autograd():
    torch.manual_seed(0)
    N,n,nt=7,12,17
    o=torch.randn(N,3)*0.3; d=torch.nn.functional.normalize(torch.randn(N,3),dim=-1)
    mid=torch.sort(torch.rand(N,n)*2.2,dim=-1)[0]; dists=torch.rand(N,n)*0.1+0.02
    sdf=(torch.randn(N,n)*0.2).double().requires_grad_(True); grad=(torch.randn(N,n,3)).double().requires_grad_(True)
    col=torch.rand(N,n,3).double().requires_grad_(True); inv_s=torch.tensor([[20.0]],dtype=torch.float64,requires_grad=True)
    bga=torch.rand(N,nt).double().mul(0.3).requires_grad_(True); bgc=torch.rand(N,nt,3).double().requires_grad_(True)
    bg_rgb=torch.ones(1,3).double(); a=0.3
    o,d,mid,dists=o.double(),d.double(),mid.double(),dists.double()
    def fwd(use_bg):
        dirs=d[:,None,:].expand(N,n,3)
        tc=(dirs*grad).sum(-1)
        ic=-(torch.relu(-tc*0.5+0.5)*(1-a)+torch.relu(-tc)*a)
        en=sdf+ic*dists*0.5; ep=sdf-ic*dists*0.5
        pc=torch.sigmoid(ep*inv_s); nc=torch.sigmoid(en*inv_s)
        alpha=((pc-nc+1e-5)/(pc+1e-5)).clip(0,1)
        pts=o[:,None,:]+d[:,None,:]*mid[...,None]; pn=pts.norm(dim=-1)
        inside=(pn<1.0).double(); relax=(pn<1.2).double()
        c=col
        if use_bg:
            alpha=alpha*inside+bga[:,:n]*(1-inside); alpha=torch.cat([alpha,bga[:,n:]],-1)
            c=col*inside[...,None]+bgc[:,:n]*(1-inside)[...,None]; c=torch.cat([c,bgc[:,n:]],1)
        w=alpha*torch.cumprod(torch.cat([torch.ones(N,1,dtype=torch.float64),1-alpha+1e-7],-1),-1)[:,:-1]
        color=(c*w[...,None]).sum(1)+bg_rgb*(1-w.sum(-1,keepdim=True))
        ge=(relax*(grad.norm(dim=-1)-1)**2).sum()/(relax.sum()+1e-5)
        return color,w,ge,dict(tc=tc,ep=ep,en=en,pc=pc,nc=nc,inside=inside,relax=relax,alpha_f=((pc-nc+1e-5)/(pc+1e-5)))
    for use_bg in (False,True):
        color,w,ge,q=fwd(use_bg)
        dcol=torch.randn(N,3).double(); dw=torch.randn(N,w.shape[1]).double(); dge=torch.tensor(0.7).double()
        L=(color*dcol).sum()+(w*dw).sum()+ge*dge
        leaves=[sdf,grad,col,inv_s]+([bga,bgc] if use_bg else [])
        ref=torch.autograd.grad(L,leaves)
        # ---- kernel algorithm (python mirror)
        ntot=w.shape[1]
        d_sdf=torch.zeros(N,n).double(); d_grad=torch.zeros(N,n,3).double(); d_cols=torch.zeros(N,n,3).double(); d_s=0.0
        d_bga=torch.zeros(N,ntot).double(); d_bgc=torch.zeros(N,ntot,3).double()
        cnt=q['relax'].sum()
        with torch.no_grad():
          for r in range(N):
            S=0.0; dbg=(dcol[r]*bg_rgb[0]).sum()
            # mixed alpha/colors
            for i in range(ntot-1,-1,-1):
                if i<n:
                    af=q['alpha_f'][r,i].clip(0,1); ins=q['inside'][r,i]
                    if use_bg: al=af*ins+bga[r,i]*(1-ins); ci=col[r,i]*ins+bgc[r,i]*(1-ins)
                    else: al=af; ci=col[r,i]
                else: al=bga[r,i]; ci=bgc[r,i]
                wi=w[r,i]; dW=(dcol[r]*ci).sum()-dbg+dw[r,i]
                om=1-al+1e-7; T=wi/al if al>0 else None
                if T is None:
                    T=1.0
                    for j in range(i):
                        if j<n:
                            afj=q['alpha_f'][r,j].clip(0,1); insj=q['inside'][r,j]
                            alj=afj*insj+bga[r,j]*(1-insj) if use_bg else afj
                        else: alj=bga[r,j]
                        T=T*(1-alj+1e-7)
                d_alpha=dW*T-S/om; S=S+dW*wi
                dcl=wi*dcol[r]
                if i<n:
                    ins=q['inside'][r,i] if use_bg else 1.0
                    if use_bg:
                        d_bga[r,i]=d_alpha*(1-ins); d_bgc[r,i]=dcl*(1-ins)
                    daf=d_alpha*ins; d_cols[r,i]=dcl*ins
                    qq=q['alpha_f'][r,i]; dq=daf if (qq>=0 and qq<=1) else 0.0
                    pc,nc=q['pc'][r,i],q['nc'][r,i]; den=pc+1e-5
                    dp=dq/den; dcc=-dq*(pc-nc+1e-5)/den**2
                    dpc=dp+dcc; dnc=-dp
                    dzp=dpc*pc*(1-pc); dzn=dnc*nc*(1-nc)
                    s=inv_s.item()
                    d_s+= (q['ep'][r,i]*dzp+q['en'][r,i]*dzn).item()
                    dep=s*dzp; den_=s*dzn
                    d_sdf[r,i]=dep+den_
                    dic=(den_-dep)*dists[r,i]*0.5
                    tc=q['tc'][r,i]; u=-tc*0.5+0.5; v=-tc
                    dtc=(0.5*(1-a)*(1.0 if u>0 else 0.0)+a*(1.0 if v>0 else 0.0))*dic
                    gn=grad[r,i].norm()
                    ek=dge/(cnt+1e-5)*q['relax'][r,i]*2*(gn-1)/gn
                    d_grad[r,i]=dtc*d[r]+ek*grad[r,i]
                else:
                    d_bga[r,i]=d_alpha; d_bgc[r,i]=dcl
        mine=[d_sdf,d_grad,d_cols,torch.tensor([[d_s]])]+([d_bga,d_bgc] if use_bg else [])
        for nm,a_,b_ in zip(['sdf','grad','col','inv_s','bga','bgc'],mine,ref):
            assert float((a_-b_).abs().max()) <= 1e-8 * max(1.0, float(b_.abs().max())), (use_bg, nm)
