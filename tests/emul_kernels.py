"""TEST INFRASTRUCTURE ONLY: torch emulation of the C-ABI launch table.

Each method states, in plain torch, exactly what the matching CUDA kernel in
druggen_b200/csrc computes (including the hand-derived second-order formulas), so that
  * on a CPU box the autograd wiring in druggen_b200/ops.py + block.py can be checked with
    gradcheck / gradgradcheck and against the oracle, and
  * on the GPU box each CUDA kernel can be unit-tested against its emulation.
It is installed with kernels._install_backend_for_tests(); the product never uses it.
"""
import torch


def _bf16r(t):
    return t.to(torch.bfloat16).to(t.dtype)


class EmulBackend:
    def __init__(self, emulate_bf16: bool = False):
        self.emulate_bf16 = emulate_bf16

    def _mm(self, a, b, prec):
        if self.emulate_bf16 and prec == "bf16":
            return _bf16r(a) @ _bf16r(b)
        if self.emulate_bf16 and prec == "bf16x3":       # hi + lo operands, three products, wide accumulation (csrc/gemm_tc.cu)
            ah, bh = _bf16r(a), _bf16r(b)
            al, bl = _bf16r(a - ah), _bf16r(b - bh)
            return (ah.double() @ bh.double() + ah.double() @ bl.double() + al.double() @ bh.double()).to(a.dtype)
        return a @ b

    # ---- dense
    def rows_gemm(self, a, w, w_is_nk, bias, relu, gate, out, prec, resid=None):
        a = a.to(w.dtype)
        r = self._mm(a, w.t() if w_is_nk else w, prec)
        if bias is not None:
            r = r + bias
        if relu:
            r = torch.relu(r)
        if gate is not None:
            r = r * (gate > 0).to(r.dtype)
        if resid is not None:
            r = r + resid
        out.copy_(r)

    def gemm_tn(self, a, b, out, accumulate, prec, colsum_a=None):
        a, b = a.to(out.dtype), b.to(out.dtype)
        r = self._mm(a.t(), b, prec)
        if accumulate:
            out.add_(r)
        else:
            out.copy_(r)
        if colsum_a is not None:
            colsum_a.add_(a.sum(0))

    def colsum(self, a, out):
        out.copy_(a.sum(0))

    def gate_mul(self, x, ref, out):
        out.copy_(x * (ref > 0).to(x.dtype))

    # ---- residual + layernorm
    @staticmethod
    def _stats(a, b, eps):
        z = a if b is None else a + b
        mu = z.mean(-1, keepdim=True)
        var = ((z - mu) ** 2).mean(-1, keepdim=True)
        r = torch.rsqrt(var + eps)
        return (z - mu) * r, r

    @staticmethod
    def _proj(w, xh):
        """P(w) = w - mean(w) - xh * mean(w * xh)  (symmetric, per row)."""
        return w - w.mean(-1, keepdim=True) - xh * (w * xh).mean(-1, keepdim=True)

    def add_ln_fwd(self, a, b, gamma, beta, out, eps):
        xh, _ = self._stats(a, b, eps)
        out.copy_(xh * gamma + beta)

    def add_ln_bwd(self, dy, a, b, gamma, dz, dgamma, dbeta, eps, accumulate=False):
        xh, r = self._stats(a, b, eps)
        val = r * self._proj(dy * gamma, xh)
        if accumulate:
            dz.add_(val)
        else:
            dz.copy_(val)
        d = a.shape[-1]
        dgamma.copy_((dy * xh).reshape(-1, d).sum(0))
        dbeta.copy_(dy.reshape(-1, d).sum(0))

    def add_ln_bwd_bwd(self, u, vg, vb, dy, a, b, gamma, g_dy, g_z, g_gamma, eps):
        xh, r = self._stats(a, b, eps)
        gh = dy * gamma
        pu = self._proj(u, xh)
        pg = self._proj(gh, xh)
        c2 = (gh * xh).mean(-1, keepdim=True)
        am = (u * pg).mean(-1, keepdim=True)
        bm = (u * xh).mean(-1, keepdim=True)
        gdy = gamma * r * pu
        gz = -(r * r) * (am * xh + c2 * pu + bm * pg)
        if vg is not None:
            gdy = gdy + vg * xh
            gz = gz + r * self._proj(vg * dy, xh)
        if vb is not None:
            gdy = gdy + vb
        g_dy.copy_(gdy)
        g_z.copy_(gz)
        g_gamma.copy_((dy * r * pu).reshape(-1, a.shape[-1]).sum(0))

    # ---- modulate
    def modulate_fwd(self, q, k, e, c, out):
        out.copy_(c * q[:, :, None, :] * k[:, None, :, :] * (e * e + e))

    def modulate_bwd(self, da, q, k, e, c, dq, dk, de):
        qi, kj = q[:, :, None, :], k[:, None, :, :]
        phi = e * e + e
        de.copy_(da * c * qi * kj * (2 * e + 1))
        dq.copy_((c * da * phi * kj).sum(2))
        dk.copy_((c * da * phi * qi).sum(1))

    def modulate_bwd_bwd(self, uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e):
        qi, kj = q[:, :, None, :], k[:, None, :, :]
        uqi, ukj = uq[:, :, None, :], uk[:, None, :, :]
        phi, dphi = e * e + e, 2 * e + 1
        mix = kj * uqi + qi * ukj
        g_da.copy_(c * (phi * mix + qi * kj * dphi * ue))
        g_e.copy_(c * da * (dphi * mix + 2 * qi * kj * ue))
        g_q.copy_((c * da * (phi * ukj + kj * dphi * ue)).sum(2))
        g_k.copy_((c * da * (phi * uqi + qi * dphi * ue)).sum(1))

    # ---- softmax over keys + aggregate
    def softmax_agg_fwd(self, a, v, out):
        out.copy_((torch.softmax(a, dim=2) * v[:, None, :, :]).sum(2))

    def softmax_agg_bwd(self, dg, a, v, da, dv, accumulate=False):
        p = torch.softmax(a, dim=2)
        vj = v[:, None, :, :]
        g = (p * vj).sum(2, keepdim=True)
        dgi = dg[:, :, None, :]
        val = p * dgi * (vj - g)
        if accumulate:
            da.add_(val)
        else:
            da.copy_(val)
        dv.copy_((p * dgi).sum(1))

    def softmax_agg_bwd_bwd(self, ua, uv, dg, a, v, g_dg, g_a, g_v):
        p = torch.softmax(a, dim=2)
        vj, uvj = v[:, None, :, :], uv[:, None, :, :]
        g = (p * vj).sum(2, keepdim=True)
        dgi = dg[:, :, None, :]
        w = ua * (vj - g) + uvj
        wbar = (p * w).sum(2, keepdim=True)
        m = (p * ua).sum(2, keepdim=True)
        g_dg.copy_(wbar.squeeze(2))
        g_a.copy_(dgi * p * (w - wbar - m * (vj - g)))
        g_v.copy_((dgi * p * (ua - m)).sum(1))

    # ---- fused
    def mlp_fwd(self, x, w1, b1, w2, b2, gamma, beta, out, eps, workspace):
        h = torch.relu(self._mm(x, w1.t(), "bf16") + b1)
        z = x + self._mm(h, w2.t(), "bf16") + b2
        self.add_ln_fwd(z, None, gamma, beta, out, eps)

    def attn_scores_fwd(self, q, k, v, e, c, a, g, stats=None):
        if a is None:
            a = torch.empty_like(e)
        self.modulate_fwd(q, k, e, c, a)
        self.softmax_agg_fwd(a, v, g)
        if stats is not None:
            m = a.max(dim=2).values
            stats[0].copy_(m)
            stats[1].copy_(1.0 / torch.exp(a - m[:, :, None, :]).sum(2))

    def attn_scores_bwd(self, dg, da_in, q, k, v, e, c, de, dq, dk, dv, stats=None, scores_bf16=False, accumulate_de=False):
        a = torch.empty_like(e)
        self.modulate_fwd(q, k, e, c, a)
        if scores_bf16:
            a = _bf16r(a)
        da = torch.empty_like(e)
        self.softmax_agg_bwd(dg, a, v, da, dv)
        if da_in is not None:
            da = da + da_in.to(da.dtype)
        de32 = torch.empty_like(e)
        self.modulate_bwd(da, q, k, e, c, dq, dk, de32)
        if accumulate_de:
            de.add_(de32)
        else:
            de.copy_(de32)

    def mlp_bwd_ln(self, x, dout, w1, b1, w2, b2, gamma, dz, h16, dgamma, dbeta, eps, workspace, mask=None):
        h = torch.relu(self._mm(x, w1.t(), "bf16") + b1)
        hq = h.to(torch.bfloat16)
        if h16 is not None:
            h16.copy_(hq)
        if mask is not None:          # bit i of word w = h[w*64+i] > 0
            bits = (h > 0).view(h.shape[0], -1, 64).to(torch.int64)
            w = torch.zeros(h.shape[0], h.shape[1] // 64, dtype=torch.int64)
            for i in range(64):
                w |= bits[:, :, i] << i         # (bit 63 wraps into the sign: same 64 bits)
            mask.copy_(w)
        m = self._mm(hq.to(x.dtype), w2.t(), "bf16") + b2
        if dgamma is None:
            dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        self.add_ln_bwd(dout, x, m, gamma, dz, dgamma, dbeta, eps)

    def mlp_bwd_dgrad(self, dz, h16, w1, w2, dx, dh16, workspace, mask=None):
        if mask is not None:
            gate = torch.stack([(mask >> i) & 1 for i in range(64)], dim=2).reshape(mask.shape[0], -1).to(dz.dtype)
        else:
            gate = (h16 > 0).to(dz.dtype)
        dh = self._mm(dz, w2, "bf16") * gate
        dhq = dh.to(torch.bfloat16)
        if dh16 is not None:
            dh16.copy_(dhq)
        dx.copy_(dz + self._mm(dhq.to(dz.dtype), w1, "bf16"))

    def softmax_agg16_fwd(self, a16, v, g, stats=None):
        a = a16.to(v.dtype).view(v.shape[0], v.shape[1], v.shape[1], v.shape[2])
        self.softmax_agg_fwd(a, v, g)
        if stats is not None:
            m = a.max(dim=2).values
            stats[0].copy_(m)
            stats[1].copy_(1.0 / torch.exp(a - m[:, :, None, :]).sum(2))

    def attn_edge_fwd(self, y, q, k, we, be, woe, boe, gamma, beta, c, out, a16, e_out, z_out, eps, workspace):
        b, n, d = q.shape
        e = self._mm(y, we.t(), "bf16") + be
        a = torch.empty_like(e).view(b, n, n, d)
        self.modulate_fwd(q, k, e.view(b, n, n, d), c, a)
        a = a.view(-1, d)
        if e_out is not None:
            e_out.copy_(e)
        if a16 is not None:
            a16.copy_(a)
        z = y + self._mm(a, woe.t(), "bf16") + boe
        if z_out is not None:
            z_out.copy_(z)
        self.add_ln_fwd(z, None, gamma, beta, out, eps)

    def symmetrize(self, e, out):
        out.copy_((e + e.permute(0, 2, 1, 3)) / 2)

    def label2onehot(self, labels, out, classes):
        out.zero_()
        out.scatter_(out.dim() - 1, labels.long().unsqueeze(-1), 1.0)

    def argmax_last(self, x, out):
        out.copy_(torch.max(x, -1)[1])

    # ---- either side of the encoder
    def embed_labels_fwd(self, labels, lut, y, n, sym):
        t = lut[labels.long()]
        y.copy_((t + t.transpose(1, 2)) / 2 if sym else t)

    def embed_labels_bwd(self, labels, dy, dlut, n, sym):
        lab = labels.long()
        d = dy.shape[-1]
        if sym:
            dlut.index_add_(0, lab.reshape(-1), dy.reshape(-1, d) * 0.5)
            dlut.index_add_(0, lab.transpose(1, 2).reshape(-1), dy.reshape(-1, d) * 0.5)
        else:
            dlut.index_add_(0, lab.reshape(-1), dy.reshape(-1, d))

    def gp_interp(self, labels, fake, eps, out, rows_per_mol):
        real = torch.nn.functional.one_hot(labels.long(), fake.shape[-1]).to(fake.dtype)
        e = eps.reshape([-1] + [1] * (fake.dim() - 1))
        out.copy_(e * real + (1 - e) * fake)

    def gp_penalty(self, g_node, g_edge, penalty, coef, scratch):
        b = g_node.shape[0]
        nrm = torch.cat([g_node.reshape(b, -1), g_edge.reshape(b, -1)], 1).norm(2, dim=1)
        penalty.copy_(((nrm - 1) ** 2).mean().reshape(1))
        coef.copy_(torch.where(nrm > 0, 2 * (nrm - 1) / (b * nrm), torch.zeros_like(nrm)))

    def gp_penalty_bwd(self, g, coef, upstream, out):
        out.copy_(upstream.reshape(()) * coef.reshape([-1] + [1] * (g.dim() - 1)) * g)

    def readout_argmax(self, x, w, bias, logits, idx):
        lg = x @ w.t() + bias
        if logits is not None:
            logits.copy_(lg)
        if idx is not None:
            idx.copy_(torch.max(lg, -1)[1])

    def adamw_flat(self, p, g, m, v, segs, nseg, lr, beta1, beta2, eps, wd):
        import numpy as np
        from druggen_b200.optim import SEG_DTYPE
        for s in np.frombuffer(segs.cpu().numpy().tobytes(), dtype=SEG_DTYPE):
            if not s["active"]:
                continue
            sl = slice(int(s["begin"]), int(s["end"]))
            p[sl].mul_(1 - lr * wd)
            m[sl].lerp_(g[sl], 1 - beta1)
            v[sl].mul_(beta2).addcmul_(g[sl], g[sl], value=1 - beta2)
            p[sl].addcdiv_(m[sl], v[sl].sqrt() / float(s["bc2_sqrt"]) + eps, value=-lr / float(s["bc1"]))

    def label_error(self, clear=True):
        return 0

    # ---- producer of the inputs / metric (restated through the oracle: csrc/glue.cu)
    def to_dense_adj(self, edge_index, batch, edge_attr, adj, cum):
        from oracle import data_oracle as orc
        b, n = adj.shape[0], adj.shape[1]
        adj.copy_(torch.from_numpy(orc.to_dense_adj(edge_index.numpy(), batch.numpy(), None if edge_attr is None else edge_attr.numpy(),
                                                    max_num_nodes=n, batch_size=b)).to(adj.dtype))

    def narrow_labels(self, adj, out, classes):
        out.copy_(adj.to(torch.uint8))

    def pack_bits(self, vecs, bits, cnt):
        import numpy as np
        v = (vecs.numpy() != 0)
        rows, f = v.shape
        w = bits.shape[1]
        pad = np.zeros((rows, w * 64), bool)
        pad[:, :f] = v
        words = (pad.reshape(rows, w, 64).astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(-1, dtype=np.uint64)
        bits.copy_(torch.from_numpy(words.view(np.int64)))
        cnt.copy_(torch.from_numpy(v.sum(1).astype(np.int32)))

    def tanimoto_agg(self, sbits, scnt, gbits, gcnt, agg, p, out_max, out_sum):
        import numpy as np
        unpack = lambda b: ((b.numpy().view(np.uint64)[:, :, None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)).reshape(b.shape[0], -1).astype(np.float32)  # noqa: E731
        xs, yg = unpack(sbits), unpack(gbits)
        tp = xs @ yg.T
        with np.errstate(invalid="ignore", divide="ignore"):
            jac = tp / (xs.sum(1, keepdims=True) + yg.sum(1)[None, :] - tp)
        jac[np.isnan(jac)] = 1
        if agg == 0:
            out_max.copy_(torch.maximum(out_max, torch.from_numpy(jac.max(0))))
        else:
            out_sum.add_(torch.from_numpy((jac.astype(np.float64) ** p).sum(0)))
