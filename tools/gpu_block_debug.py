"""Block-level parity on GPU: engine._resblock / _transformer / time_embeddings vs the oracle (debug aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import svd_oracle as O  # noqa: E402
from tests.common import SVD, TINY, build_models, make_inputs, oracle_cfg, rel_l2, state  # noqa: E402

kind = SVD if (len(sys.argv) > 1 and sys.argv[1] == "svd") else TINY
cfg = oracle_cfg(kind)
unet, _ = build_models(kind, controlnet=False)
usd = state(unet)
B, F, h, w = 2, 14, 16, 24
sample, ehs, ati, cond = make_inputs(B, F, h, w)
dev = "cuda"
unet.to(dev)
eng = unet._get_engine()
t = torch.tensor(1.63777)
C0 = kind["block_out_channels"][0]
with torch.no_grad():
    # --- time embeddings
    emb = O._embed(usd, cfg, t, ati, B, torch.float32, "cpu")
    temb = eng.time_embeddings(t.expand(B).contiguous().to(dev), ati.to(dev))
    r0 = eng.down[0]["res"][0]
    ref_s = O.linear(usd, "down_blocks.0.resnets.0.spatial_res_block.time_emb_proj", torch.nn.functional.silu(emb))
    print("temb spatial slice:", rel_l2(temb[:, r0.temb_off_s:r0.temb_off_s + C0], ref_s))
    # --- pos emb + kv
    eng._ensure_pos_emb(F)
    kvs = eng.context_kv(ehs.to(dev))
    t0 = eng.down[0]["tf"][0]
    frames = torch.arange(F)
    pe = O.timestep_embedding(usd, "down_blocks.0.attentions.0.time_pos_embed", O.timesteps_sinusoid(frames, C0))
    print("pos_emb:", rel_l2(t0.pos_emb, pe))
    kref = O.linear(usd, "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k", ehs)
    print("ctx k:", rel_l2(kvs[0][0].view(B, -1, C0), kref))
    # --- resblock
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B * F, C0, h, w, generator=g)
    ind = torch.zeros(B, F)
    embf = emb.repeat_interleave(F, 0)
    ref = O.spatio_temporal_res_block(usd, "down_blocks.0.resnets.0", x, embf, ind, 1e-6)
    xt = x.permute(0, 2, 3, 1).reshape(-1, C0).to(dev, torch.bfloat16).contiguous()
    out = eng._resblock(r0, xt, None, B=B, F=F, H=h, W=w, temb=temb)
    torch.cuda.synchronize()
    o = out.view(B * F, h, w, C0).permute(0, 3, 1, 2)
    print("resblock:", rel_l2(o, ref), "finite:", bool(torch.isfinite(out.float()).all()))
    # spatial part only
    refs = O.resnet_block_2d(usd, "down_blocks.0.resnets.0.spatial_res_block", x, embf, 1e-6)
    # --- transformer
    ehs_bf = ehs.repeat_interleave(F, 0)
    reft = O.transformer_spatio_temporal(usd, "down_blocks.0.attentions.0", x, ehs_bf, ind, kind["num_attention_heads"][0])
    outt = eng._transformer(t0, xt, kvs[0], B=B, F=F, H=h, W=w, n_ctx=B, batch_offset=0)
    torch.cuda.synchronize()
    ot = outt.view(B * F, h, w, C0).permute(0, 3, 1, 2)
    print("transformer:", rel_l2(ot, reft), "finite:", bool(torch.isfinite(outt.float()).all()))
    # --- step by step inside the transformer
    S = h * w
    rows = B * F * S
    y = eng._gn(xt, t0.gn_g, t0.gn_b, rows=rows, rows_per_inst=S, eps=1e-6, silu=False)
    p = "down_blocks.0.attentions.0"
    hh = torch.nn.functional.group_norm(x, 32, usd[p + ".norm.weight"], usd[p + ".norm.bias"], 1e-6)
    hh = hh.permute(0, 2, 3, 1).reshape(B * F, S, C0)
    print(" gn:", rel_l2(y.view(B * F, S, C0), hh))
    hh = O.linear(usd, p + ".proj_in", hh)
    hg = eng._linear(y, t0.w_in, M=rows, bias=t0.b_in)
    print(" proj_in:", rel_l2(hg.view(B * F, S, C0), hh))
    sb = p + ".transformer_blocks.0"
    heads = kind["num_attention_heads"][0]
    a1 = O.attention(usd, sb + ".attn1", O.layer_norm(usd, sb + ".norm1", hh), None, heads)
    yg = eng._ln(hg, t0.ln["s1"], rows=rows, C=C0)
    print(" ln:", rel_l2(yg.view(B * F, S, C0), O.layer_norm(usd, sb + ".norm1", hh)))
    qkv = eng._linear(yg, t0.s_attn1.wqkv, M=rows)
    from this_and_that_vdm_b200 import lib
    og = torch.empty_like(yg)
    lib.attn_spatial(qkv, qkv[:, C0:], qkv[:, 2 * C0:], og, ldq=3 * C0, ldk=3 * C0, ldv=3 * C0, ldo=C0, n_img=B * F, heads=heads, seq=S, scale=0.125)
    a1g = eng._linear(og, t0.s_attn1.wo, M=rows, bias=t0.s_attn1.bo)
    print(" attn1:", rel_l2(a1g.view(B * F, S, C0), a1))
    hh2 = hh + a1
    a2 = O.attention(usd, sb + ".attn2", O.layer_norm(usd, sb + ".norm2", hh2), ehs_bf, heads)
    h2g = (hh2.reshape(rows, C0)).to(dev, torch.bfloat16).contiguous()
    yg = eng._ln(h2g, t0.ln["s2"], rows=rows, C=C0)
    q = eng._linear(yg, t0.s_attn2.wq, M=rows)
    ks, vs, kt, vt = kvs[0]
    L = ks.shape[0] // B
    lib.attn_cross(q, ks, vs, og, ldq=C0, ldo=C0, rows=rows, heads=heads, L=L, F=F, S=S, n_ctx=B, temporal=False, batch_offset=0, scale=0.125)
    a2g = eng._linear(og, t0.s_attn2.wo, M=rows, bias=t0.s_attn2.bo)
    print(" attn2:", rel_l2(a2g.view(B * F, S, C0), a2))
    hh3 = hh2 + a2
    f3 = O.feed_forward(usd, sb + ".ff", O.layer_norm(usd, sb + ".norm3", hh3))
    h3g = hh3.reshape(rows, C0).to(dev, torch.bfloat16).contiguous()
    yg = eng._ln(h3g, t0.ln["s3"], rows=rows, C=C0)
    gg = eng._linear(yg, t0.s_ff.w1, M=rows, bias=t0.s_ff.b1, geglu=True)
    f3g = eng._linear(gg, t0.s_ff.w2, M=rows, bias=t0.s_ff.b2)
    print(" ff:", rel_l2(f3g.view(B * F, S, C0), f3))
    hs = hh3 + f3
    # temporal
    tb = p + ".temporal_transformer_blocks.0"
    hsg = hs.reshape(rows, C0).to(dev, torch.bfloat16).contiguous()
    hm = torch.empty_like(hsg)
    yg = eng._ln(hsg, t0.ln["tin"], rows=rows, C=C0, addvec=t0.pos_emb, F=F, S=S, sum_out=hm)
    emb_pos = pe.repeat(B, 1)[:, None, :]
    hmix = hs + emb_pos
    print(" hm:", rel_l2(hm.view(B * F, S, C0), hmix))
    xt_ = hmix.reshape(B, F, S, C0).permute(0, 2, 1, 3).reshape(B * S, F, C0)
    a = O.attention(usd, tb + ".attn1", O.layer_norm(usd, tb + ".norm1", xt_), None, heads)
    xg = xt_.reshape(B, S, F, C0).permute(0, 2, 1, 3).reshape(rows, C0).to(dev, torch.bfloat16).contiguous()
    yg = eng._ln(xg, t0.ln["t1"], rows=rows, C=C0)
    qkv = eng._linear(yg, t0.t_attn1.wqkv, M=rows)
    lib.attn_temporal(qkv, qkv[:, C0:], qkv[:, 2 * C0:], og, ldq=3 * C0, ldk=3 * C0, ldv=3 * C0, ldo=C0, B=B, F=F, S=S, heads=heads, scale=0.125)
    ag = eng._linear(og, t0.t_attn1.wo, M=rows, bias=t0.t_attn1.bo)
    aref = a.reshape(B, S, F, C0).permute(0, 2, 1, 3).reshape(B * F, S, C0)
    print(" t_attn1:", rel_l2(ag.view(B * F, S, C0), aref))
    first = ehs_bf.reshape(B, F, -1, 1024)[:, 0]
    tc = first[None].broadcast_to(S, B, first.shape[1], 1024).reshape(S * B, first.shape[1], 1024)
    a = O.attention(usd, tb + ".attn2", O.layer_norm(usd, tb + ".norm2", xt_), tc, heads)
    yg = eng._ln(xg, t0.ln["t2"], rows=rows, C=C0)
    q = eng._linear(yg, t0.t_attn2.wq, M=rows)
    lib.attn_cross(q, kt, vt, og, ldq=C0, ldo=C0, rows=rows, heads=heads, L=L, F=F, S=S, n_ctx=B, temporal=True, batch_offset=0, scale=0.125)
    ag = eng._linear(og, t0.t_attn2.wo, M=rows, bias=t0.t_attn2.bo)
    aref = a.reshape(B, S, F, C0).permute(0, 2, 1, 3).reshape(B * F, S, C0)
    print(" t_attn2:", rel_l2(ag.view(B * F, S, C0), aref))
