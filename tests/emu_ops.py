"""CPU emulation of every kernel contract of ``vrdone_b200.cuda_ops.CudaOps``  --  TEST INFRASTRUCTURE.

Each method is the executable specification of one CUDA kernel, written with plain (slow, per-sequence)
torch code on CPU tensors.  Tests use it (a) on the CPU to check that the engine's varlen schedule
reproduces the padded reference forward, and (b) on the GPU box as the per-kernel parity reference.
The product never imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPS = 1e-5


def _ln_rows(x, g, b):
    x = x.float()
    mu = x.mean(-1, keepdim=True)
    r = x - mu
    var = (r * r).mean(-1, keepdim=True)
    return r / torch.sqrt(var + EPS) * g + b


class EmuOps:
    def __init__(self):
        self.calls = []

    # helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _seqs(lay, streams=1):
        for s in range(streams):
            for i in range(lay.B):
                yield s, i, s * lay.R + int(lay.off[i]), int(lay.len[i]), int(lay.haspad[i])

    @staticmethod
    def _valid_rows(lay, streams):
        v = (lay.row_seq.cpu() >= 0)
        return v.repeat(streams)

    # kernels ---------------------------------------------------------------------------------
    def pack_pairs(self, feats, strides, lay, nv, nc, nbs, nbe, vis, clp, bso, bent, token_major=False):
        """``feats``: in the emulation, the list of (C, L) tensors themselves (the CUDA op takes a pointer table)."""
        self.calls.append("pack_pairs")
        R = lay.R
        vis.zero_(); bso.zero_(); bent.zero_()
        if clp is not None:
            clp.zero_()
        c0 = 2 * nv + 2 * nc
        for _, i, r0, L, _ in self._seqs(lay):
            f = feats[i].t().float()   # (L, C)
            assert f.shape[0] == L
            vis[r0:r0 + L] = f[:, :nv].to(vis.dtype)
            vis[R + r0:R + r0 + L] = f[:, nv:2 * nv].to(vis.dtype)
            if clp is not None:
                clp[r0:r0 + L] = f[:, 2 * nv:2 * nv + nc].to(clp.dtype)
                clp[R + r0:R + r0 + L] = f[:, 2 * nv + nc:c0].to(clp.dtype)
            bso[r0:r0 + L, :nbs] = f[:, c0:c0 + nbs]
            bent[r0:r0 + L, :nbe] = f[:, c0 + nbs:c0 + nbs + nbe]
            bent[R + r0:R + r0 + L, :nbe] = f[:, c0 + nbs + nbe:c0 + nbs + 2 * nbe]

    def gemm(self, a, w, out, bias=None, taps=1, act=0, res1=None, res2=None, corr=None, lay=None, streams=1):
        self.calls.append("gemm")
        M, K = a.shape
        af = a.float()
        if taps == 3:
            z = torch.zeros(1, K)
            af = torch.cat([torch.cat([z, af[:-1]]), af, torch.cat([af[1:], z])], 1)   # rows r-1, r, r+1
        acc = af @ w.float().t()
        if bias is not None:
            acc = acc + bias
        if corr is not None:
            for s, i, r0, L, hp in self._seqs(lay, streams):
                if hp:
                    acc[r0 + L - 1] += corr
        if act == 1:
            acc = F.relu(acc)
        elif act == 2:
            acc = F.gelu(acc)
        if res1 is not None:
            acc = acc + res1
        if res2 is not None:
            acc = acc + res2
        if lay is not None:
            acc = acc * self._valid_rows(lay, streams)[:, None]
        out.copy_(acc.to(out.dtype))

    def gemm_ln(self, a, w, out, ln, bias=None, taps=1, corr=None, relu=False, lay=None, streams=1):
        """Conv-as-GEMM with the channel LayerNorm (+ ReLU) as its epilogue: the fp32 accumulator is normalised, never stored."""
        self.calls.append("gemm_ln")
        M, K = a.shape
        af = a.float()
        if taps == 3:
            z = torch.zeros(1, K)
            af = torch.cat([torch.cat([z, af[:-1]]), af, torch.cat([af[1:], z])], 1)   # rows r-1, r, r+1
        acc = af @ w.float().t()
        if bias is not None:
            acc = acc + bias
        if corr is not None:
            for s, i, r0, L, hp in self._seqs(lay, streams):
                if hp:
                    acc[r0 + L - 1] += corr
        y = _ln_rows(acc, ln[0], ln[1])
        if relu:
            y = F.relu(y)
        if lay is not None:
            y = y * self._valid_rows(lay, streams)[:, None]
        out.copy_(y.to(out.dtype))

    def gemm_res_ln(self, a, w, out, ln_out, ln, bias=None, res1=None, lay=None, streams=1):
        """Projection + residual (fp32 out) and the LayerNorm of the sum (bf16 ln_out)."""
        self.calls.append("gemm_res_ln")
        acc = a.float() @ w.float().t()
        if bias is not None:
            acc = acc + bias
        acc = acc + res1
        if lay is not None:
            acc = acc * self._valid_rows(lay, streams)[:, None]
        out.copy_(acc)
        y = _ln_rows(acc, ln[0], ln[1])
        if lay is not None:
            y = y * self._valid_rows(lay, streams)[:, None]
        ln_out.copy_(y.to(ln_out.dtype))

    def layernorm(self, x, g, b, out, relu=False, lay=None, streams=1):
        self.calls.append("layernorm")
        y = _ln_rows(x, g, b)
        if relu:
            y = F.relu(y)
        if lay is not None:
            y = y * self._valid_rows(lay, streams)[:, None]
        out.copy_(y.to(out.dtype))

    def small_conv(self, x, cin, w, bias, ln, relu, out, lay, streams):
        """k=3 conv with tiny cin on fp32 rows (zero separators give the zero padding) + bias [+ LN] [+ ReLU]."""
        self.calls.append("small_conv")
        xf = x[:, :cin].float()
        z = torch.zeros(1, cin)
        a = torch.cat([torch.cat([z, xf[:-1]]), xf, torch.cat([xf[1:], z])], 1)
        y = a @ w + bias          # w: [3*cin, N]
        if ln is not None:
            y = _ln_rows(y, ln[0], ln[1])
        if relu:
            y = F.relu(y)
        y = y * self._valid_rows(lay, streams)[:, None]
        out.copy_(y.to(out.dtype))

    def dwconv_ln(self, x, lay_in, lay_out, stride, pre, branches, streams):
        """For every valid output row: depthwise k=3 (stride 1|2) over [LN_pre](x) with the reference's padded-batch
        edge semantics, then a LayerNorm per branch.  Left edge: zero.  Right neighbour == first pad column: LN_pre bias
        for pre-normalised branches (0 for raw ones) if the pad column exists, else zero (stride 2 always has it)."""
        self.calls.append("dwconv_ln")
        xf = x.float()
        C = xf.shape[1]
        for w, use_pre, g, b, out in branches:
            out.zero_()
        for (s, i, r_in, L_in, hp), (_, _, r_out, L_out, _) in zip(self._seqs(lay_in, streams), self._seqs(lay_out, streams)):
            seq = xf[r_in:r_in + L_in]
            for w, use_pre, g, b, out in branches:
                if use_pre:
                    src = _ln_rows(seq, pre[0], pre[1])
                    padv = pre[1] if (hp or stride == 2) else torch.zeros(C)
                else:
                    src = seq
                    padv = torch.zeros(C)
                ext = torch.cat([torch.zeros(1, C), src, padv[None], torch.zeros(1, C)])   # index t+1 <-> time t
                t = torch.arange(L_out) * stride
                y = ext[t] * w[0] + ext[t + 1] * w[1] + ext[t + 2] * w[2]      # w: [3, C]
                out[r_out:r_out + L_out] = _ln_rows(y, g, b).to(out.dtype)

    def window_attn(self, q, k, v, out, lay, n_head, w, streams):
        self.calls.append("window_attn")
        out.zero_()
        C = q.shape[1]
        hs = C // n_head
        for s, i, r0, L, _ in self._seqs(lay, streams):
            qq = q[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            kk = k[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            vv = v[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            att = qq @ kk.transpose(1, 2)
            idx = torch.arange(L)
            band = (idx[:, None] - idx[None, :]).abs() <= w
            att = att.masked_fill(~band[None], float("-inf"))
            o = F.softmax(att, -1) @ vv
            out[r0:r0 + L] = o.transpose(0, 1).reshape(L, C).to(out.dtype)

    def full_attn(self, q, k, v, out, lay, n_head):
        self.calls.append("full_attn")
        out.zero_()
        C = q.shape[1]
        hs = C // n_head
        for s, i, r0, L, _ in self._seqs(lay, 1):
            qq = q[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            kk = k[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            vv = v[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            o = F.softmax(qq @ kk.transpose(1, 2), -1) @ vv
            out[r0:r0 + L] = o.transpose(0, 1).reshape(L, C).to(out.dtype)

    def maxpool_skip(self, x, lay_in, lay_out, out):
        """out[i] = max(x[2i-1], x[2i], x[2i+1]); index -1 is ignored, index L (a zero pad column, it exists whenever
        it is read) counts as 0."""
        self.calls.append("maxpool_skip")
        out.zero_()
        C = x.shape[1]
        for (s, i, r_in, L_in, _), (_, _, r_out, L_out, _) in zip(self._seqs(lay_in), self._seqs(lay_out)):
            seq = x[r_in:r_in + L_in].float()
            ext = torch.cat([torch.full((1, C), float("-inf")), seq, torch.zeros(1, C)])
            t = torch.arange(L_out) * 2
            out[r_out:r_out + L_out] = torch.maximum(torch.maximum(ext[t], ext[t + 1]), ext[t + 2])

    def fpn_top(self, x, lay, pre, w, ln, out):
        """Grouped k=3 conv (F groups, 2 input channels each) over LN_pre(x), then LN."""
        self.calls.append("fpn_top")
        out.zero_()
        Fd = w.shape[1] // 2          # w: [3, 2*F], indexed by input channel 2c + j
        for s, i, r0, L, hp in self._seqs(lay):
            src = _ln_rows(x[r0:r0 + L], pre[0], pre[1])
            C = src.shape[1]
            padv = pre[1] if hp else torch.zeros(C)
            ext = torch.cat([torch.zeros(1, C), src, padv[None]])
            y = sum((ext[k:k + L] * w[k]).view(L, Fd, 2).sum(-1) for k in range(3))
            out[r0:r0 + L] = _ln_rows(y, ln[0], ln[1])

    def fpn_level(self, cur, y_up, lay, lay_up, ln_lat, beta_up, w, ln, out):
        """z = LN_lat(cur) + nearest_up2(y_up); out = LN(depthwise_k3(z)).  The first pad column of z (if it exists) is
        beta_lat + (y_up[last valid] if L odd else beta_fpn_up)."""
        self.calls.append("fpn_level")
        out.zero_()
        Fd = cur.shape[1]
        for (s, i, r0, L, hp), (_, _, ru, Lu, _) in zip(self._seqs(lay), self._seqs(lay_up)):
            up = y_up[ru:ru + Lu].float()
            z = _ln_rows(cur[r0:r0 + L], ln_lat[0], ln_lat[1]) + up[torch.arange(L) // 2]
            if hp:
                padv = ln_lat[1] + (up[Lu - 1] if L % 2 == 1 else beta_up)
            else:
                padv = torch.zeros(Fd)
            ext = torch.cat([torch.zeros(1, Fd), z, padv[None]])
            y = ext[0:L] * w[0] + ext[1:L + 1] * w[1] + ext[2:L + 2] * w[2]
            out[r0:r0 + L] = _ln_rows(y, ln[0], ln[1])

    def mask_features(self, y, lay, beta, w, bias, out):
        self.calls.append("mask_features")
        out.zero_()
        Fd = y.shape[1]
        for s, i, r0, L, hp in self._seqs(lay):
            padv = beta if hp else torch.zeros(Fd)
            ext = torch.cat([torch.zeros(1, Fd), y[r0:r0 + L].float(), padv[None]])
            out[r0:r0 + L] = (ext[0:L] * w[0] + ext[1:L + 1] * w[1] + ext[2:L + 2] * w[2] + bias).to(out.dtype)

    def query_ln(self, x, ln, pos, Q, nrows, dw, ln2, out):
        """out = LN2(dw * (LN(x) + pos[row % Q])) with every stage optional; rows >= nrows are zeroed."""
        self.calls.append("query_ln")
        y = x[:nrows].float()
        if ln is not None:
            y = _ln_rows(y, ln[0], ln[1])
        if pos is not None:
            y = y + pos[torch.arange(nrows) % Q]
        if dw is not None:
            y = y * dw.flatten()
        if ln2 is not None:
            y = _ln_rows(y, ln2[0], ln2[1])
        out.zero_()
        out[:nrows] = y.to(out.dtype)

    def query_self_attn(self, q, k, v, out, B, Q, n_head):
        self.calls.append("query_self_attn")
        D = q.shape[1]
        hs = D // n_head
        n = B * Q
        qq = q[:n].float().view(B, Q, n_head, hs).transpose(1, 2)
        kk = k[:n].float().view(B, Q, n_head, hs).transpose(1, 2)
        vv = v[:n].float().view(B, Q, n_head, hs).transpose(1, 2)
        o = F.softmax(qq @ kk.transpose(-1, -2), -1) @ vv
        out.zero_()
        out[:n] = o.transpose(1, 2).reshape(n, D).to(out.dtype)

    def query_cross_attn(self, q, k, v, out, lay, Q, n_head):
        self.calls.append("query_cross_attn")
        D = q.shape[1]
        hs = D // n_head
        out.zero_()
        for s, i, r0, L, _ in self._seqs(lay):
            qq = q[i * Q:(i + 1) * Q].float().view(Q, n_head, hs).transpose(0, 1)
            kk = k[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            vv = v[r0:r0 + L].float().view(L, n_head, hs).transpose(0, 1)
            o = F.softmax(qq @ kk.transpose(1, 2), -1) @ vv
            out[i * Q:(i + 1) * Q] = o.transpose(0, 1).reshape(Q, D).to(out.dtype)

    def mask_logits(self, me, mf, lay, Q, masks, first_last):
        """masks[row, q] = <me[pair*Q + q], mf[row]>; first_last[pair, q] = first / last t with sigmoid(logit) > 0.5 in
        fp32 (or -1, -1)."""
        self.calls.append("mask_logits")
        if masks is not None:
            masks.zero_()
        for s, i, r0, L, _ in self._seqs(lay):
            m = mf[r0:r0 + L].float() @ me[i * Q:(i + 1) * Q].float().t()      # (L, Q)
            if masks is not None:
                masks[r0:r0 + L] = m
            act = torch.sigmoid(m) > 0.5
            for qi in range(Q):
                nz = torch.nonzero(act[:, qi]).flatten()
                first_last[i, qi, 0] = int(nz[0]) if nz.numel() else -1
                first_last[i, qi, 1] = int(nz[-1]) if nz.numel() else -1

    def softmax_topk(self, logits, nrows, n_cls, topk, scores, ids):
        """softmax over the n_cls logits of a row, then the top-k of classes 1..n_cls-1 (ids are 1-based class ids;
        ties resolve to the lower id)."""
        self.calls.append("softmax_topk")
        p = F.softmax(logits[:nrows, :n_cls].float(), -1)[:, 1:]
        sc, ix = torch.sort(p, dim=-1, descending=True, stable=True)
        scores.copy_(sc[:, :topk])
        ids.copy_((ix[:, :topk] + 1).to(ids.dtype))
