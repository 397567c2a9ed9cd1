import torch, torch.nn as nn, sys
sys.path.insert(0, '.')
from multimodal_lipread_b200 import kernels as K
def err(a,b): 
    a=a.detach().cpu().double(); b=b.detach().cpu().double(); return (a-b).abs().max().item()/(b.abs().max().item()+1e-30)
for (H,I,B,T) in [(32,20,5,4),(128,576,6,29)]:
    torch.manual_seed(H)
    lstm = nn.LSTM(I, H, 1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, I, requires_grad=True)
    out,_ = lstm(x); dout = torch.randn_like(out); out.backward(dout)
    xd = x.detach().cuda()
    for d, sfx in enumerate(("", "_reverse")):
        wih, whh = getattr(lstm,"weight_ih_l0"+sfx).detach().cuda(), getattr(lstm,"weight_hh_l0"+sfx).detach().cuda()
        bih, bhh = getattr(lstm,"bias_ih_l0"+sfx).detach().cuda(), getattr(lstm,"bias_hh_l0"+sfx).detach().cuda()
        xproj = torch.empty(B*T,4*H,device="cuda"); K.linear_fwd(xd.view(B*T,I), wih, xproj, bias=bih)
        o = torch.zeros(B,T,2*H,device="cuda")
        gates,cst,hprev = torch.empty(B,T,4*H,device="cuda"),torch.empty(B,T,H,device="cuda"),torch.empty(B,T,H,device="cuda")
        K.lstm_fwd(xproj,4*H,bhh,whh,o[:,:,d*H:],2*H,gates,cst,hprev,B,T,H,T,d)
        print(H,d,"fwd",err(o[:,:,d*H:(d+1)*H], out[:,:,d*H:(d+1)*H]))
        dg = torch.zeros(B,T,4*H,device="cuda"); dod = dout.cuda()
        K.lstm_bwd(dod[:,:,d*H:],2*H,-1,gates,cst,whh,dg,B,T,H,T,d)
        db = torch.zeros(4*H,device="cuda"); K.colsum(dg,4*H,B*T,4*H,db)
        print(H,d,"db",err(db, getattr(lstm,"bias_ih_l0"+sfx).grad))
        ref_dwih = dg.view(B*T,4*H).t() @ xd.view(B*T,I)
        print(H,d,"dwih(torch matmul of our dg)",err(ref_dwih, getattr(lstm,"weight_ih_l0"+sfx).grad))
        dwih = torch.zeros_like(wih); K.linear_wgrad(dg.view(B*T,4*H), xd.view(B*T,I), dwih)
        print(H,d,"dwih(ours)",err(dwih, getattr(lstm,"weight_ih_l0"+sfx).grad), "vs matmul", err(dwih, ref_dwih))
        dwhh = torch.zeros_like(whh); K.linear_wgrad(dg.view(B*T,4*H), hprev.view(B*T,H), dwhh)
        print(H,d,"dwhh",err(dwhh, getattr(lstm,"weight_hh_l0"+sfx).grad))
