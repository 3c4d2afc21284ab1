import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
ah = float(DDPM().alpha_hat[500])
x = math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(32, 2048, 3, generator=g)
st = torch2sparse(x.to(dev))
geo = spvd_b200.Geometry(st.C.float(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1,), devox_strides=(1,), n_batch=32)
for s in (4, 2, 1):
    lv = geo.level(s); km = lv.kmap3; n = lv.n
    cin, cout, cin2 = 192, 128, 320
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    w2 = torch.randn(1, cin2, cout, device=dev) / cin2 ** 0.5
    wp, wp2 = ops.pack_weights_f16(w), ops.pack_weights_f16(w2)
    xin = torch.randn(n, cin, device=dev).half()
    xin2 = torch.randn(n, cin2, device=dev).half()
    ms, tm, perm = ops.kmap_sort(km.map_t, n, 27)
    ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, None, ops.CONV_UMMA, mt=3)
    ref2 = ref + ops.conv_fwd(xin2, w2, wp2, None, None, n, None, ops.CONV_UMMA)
    for name, kw in (("sorted", dict(map_t=ms, tile_mask=tm, out_rows=perm)), ("lex", dict(map_t=km.map_t, tile_mask=km.mask)), ("lex-nomask", dict(map_t=km.map_t))):
        for bn in (128, 64, 32):
            for st_ in (0, 2, 3) * 6:
                o32, _ = ops.conv_sp(xin, wp, 27, cin, cout, n, bn=bn, stages=st_, **kw)
                e = float((o32 - ref).norm() / ref.norm())
                bad = ((o32 - ref).abs().amax(1) > 1e-3 * ref.abs().max()).nonzero().flatten()
                o32b, _ = ops.conv_sp(xin, wp, 27, cin, cout, n, bn=bn, stages=st_, in2=xin2, w2_packed=wp2, **kw)
                e2 = float((o32b - ref2).norm() / ref2.norm())
                if e > 1e-5 or e2 > 1e-5 or False: print(f"s={s} {name} bn={bn} stages={st_}: plain err={e:.1e} bad rows={bad.numel()} first={bad[:6].tolist()} | +in2 err={e2:.1e}")

print("done")
