"""Varlen row layout of a batch of subject-object pairs in HBM.

All activations are token-major matrices ``[rows, channels]`` (channels contiguous).  At pyramid level
``l`` a pair with ``L`` valid frames owns ``L_l = ceil(L / 2**l)`` consecutive rows; a single all-zero
*separator* row precedes the first pair and follows every pair, so that the three row-shifted K-slabs
of a k=3 convolution-as-GEMM read zeros across sequence boundaries.  ``rows`` is rounded up to a
multiple of 128 (the GEMM M tile); the tail rows are separators too.

The reference pads instead (models/maskvrd.py:363-414): short pairs to ``max_seq_len``, long pairs to the
longest pair of their 200-pair slice rounded up to ``max_div_factor``.  Padding is not neutral there
(SURVEY.md section 7 hard-part 2 / appendix B): the first pad column of a level carries a per-channel constant that
k=3 convolutions read.  ``haspad[l]`` records, per pair, whether such a column exists at level ``l``
(``L_l < T_pad / 2**l``), and the kernels add the constant analytically.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

ROW_TILE = 128


def max_div_factor(mc: dict) -> int:
    """Largest fpn_stride * 2 * (win // 2) over the pyramid levels (reference maskvrd.py:57-63)."""
    n_levels = mc["backbone_arch"][-1] + 1
    w = mc["n_mha_win_size"]
    best = 1
    for l in range(mc["fpn_start_level"], n_levels):
        s = mc["scale_factor"] ** l
        best = max(best, s * (w // 2) * 2 if w > 1 else s)
    return best


def reference_padded_lengths(lengths: Sequence[int], mc: dict) -> List[int]:
    """T_pad the reference would give each pair (it decides per slice of ``max_so_pair`` pairs)."""
    msl, mdf, chunk = mc["max_seq_len"], max_div_factor(mc), mc["max_so_pair"]
    out: List[int] = []
    for s in range(0, len(lengths), chunk):
        sl = lengths[s:s + chunk]
        longest = max([msl] + [l for l in sl if l > msl])
        t_long = (longest + mdf - 1) // mdf * mdf
        out += [msl if l <= msl else t_long for l in sl]
    return out


class LevelLayout:
    """Rows of one pyramid level.  ``seqinfo[i] = (row offset, valid length, haspad, 0)`` and ``row_seq[r]`` =
    owning pair or -1 for a separator row; both live on the device as int32."""

    def __init__(self, level: int, off: np.ndarray, length: np.ndarray, haspad: np.ndarray, rows: int):
        self.level = level
        self.off, self.len, self.haspad, self.R = off, length, haspad, rows
        self.max_len = int(length.max())
        self.row_seq: torch.Tensor = None   # int32 [R]
        self.seqinfo: torch.Tensor = None   # int32 [B, 4]
        # level 0 only: the query tiles of the full-attention kernel, (first row, pair's first row, pair length, 0) per 128-row tile
        # of every pair, sorted by pair length (longest first: a static round-robin over this list balances the persistent CTAs)
        self.tiles: torch.Tensor = None     # int32 [n_tiles, 4]
        self.n_tiles = 0

    @property
    def B(self) -> int:
        return len(self.len)


class PackLayout:
    """Host-side construction of the per-level arrays; ``device`` given -> uploaded at once (one pinned staging copy);
    ``device=None`` -> the caller uploads ``host_words()`` itself (several chunks in one copy) and calls ``bind``."""

    def __init__(self, lengths: Sequence[int], tpads: Sequence[int], n_levels: int, device=None):
        lens = np.asarray(lengths, dtype=np.int64)
        tp = np.asarray(tpads, dtype=np.int64)
        assert lens.ndim == 1 and lens.shape == tp.shape and (lens >= 1).all() and (lens <= tp).all()
        assert (tp % (1 << (n_levels - 1)) == 0).all()
        self.lengths, self.tpads, self.n_levels = lens, tp, n_levels
        self.B = len(lens)
        self.levels: List[LevelLayout] = []
        host = []
        ids = np.arange(self.B, dtype=np.int32)
        for l in range(n_levels):
            ll = (lens + (1 << l) - 1) >> l
            off = np.empty(self.B, dtype=np.int64)
            off[0] = 1
            np.cumsum(ll[:-1] + 1, out=off[1:])
            off[1:] += 1
            used = int(off[-1] + ll[-1] + 1)
            rows = (used + ROW_TILE - 1) // ROW_TILE * ROW_TILE
            haspad = (ll < (tp >> l)).astype(np.int32)
            # rows: separator, then per pair its ll rows and one separator; the tail up to ``rows`` are separators too
            row_seq = np.full(rows, -1, dtype=np.int32)
            body = row_seq[1:used]
            body[:] = np.repeat(ids, ll + 1)
            body[off + ll - 1] = -1
            info = np.zeros((self.B, 4), dtype=np.int32)
            info[:, 0], info[:, 1], info[:, 2] = off, ll, haspad
            lev = LevelLayout(l, off.astype(np.int32), ll.astype(np.int32), haspad, rows)
            self.levels.append(lev)
            host += [row_seq, info.reshape(-1)]
        host.append(self._attention_tiles(self.levels[0]))
        self.levels[0].n_tiles = host[-1].size // 4
        self._host = host
        self.n_words = sum(h.size for h in host)      # int32 words; every piece is a multiple of 4 words (16 bytes)
        self.total_frames = int(lens.sum())
        if device is not None:
            if torch.device(device).type == "cuda":
                # pinned staging: a pageable source would make the copy wait for the stream's earlier work
                flat = torch.empty(self.n_words, dtype=torch.int32, pin_memory=True)
                self.host_words(flat.numpy())
                self.bind(flat.to(device, non_blocking=True))
            else:
                self.bind(torch.from_numpy(np.concatenate(host)))

    ATTN_TILE = 128

    @classmethod
    def _attention_tiles(cls, lev: "LevelLayout") -> np.ndarray:
        """int32 [n_tiles * 4]: (first layout row of the tile, pair's first row, pair length, 0), longest pairs first."""
        ln, off = lev.len.astype(np.int64), lev.off.astype(np.int64)
        n_t = (ln + cls.ATTN_TILE - 1) // cls.ATTN_TILE
        pair = np.repeat(np.arange(len(ln)), n_t)
        t_in_pair = np.arange(int(n_t.sum())) - np.repeat(np.cumsum(n_t) - n_t, n_t)
        order = np.argsort(-ln[pair], kind="stable")
        tiles = np.zeros((len(pair), 4), dtype=np.int32)
        tiles[:, 0] = (off[pair] + cls.ATTN_TILE * t_in_pair)[order]
        tiles[:, 1] = off[pair][order]
        tiles[:, 2] = ln[pair][order]
        return tiles.reshape(-1)

    def host_words(self, out: np.ndarray) -> None:
        """Write the int32 image of all levels into ``out`` (n_words elements)."""
        np.concatenate(self._host, out=out)

    def bind(self, dev: torch.Tensor) -> None:
        """``dev``: int32 tensor of n_words elements (16-byte aligned) holding the image written by ``host_words``."""
        pos = 0
        for lev in self.levels:
            lev.row_seq = dev[pos:pos + lev.R]
            pos += lev.R
            lev.seqinfo = dev[pos:pos + 4 * self.B].view(self.B, 4)
            pos += 4 * self.B
        l0 = self.levels[0]
        l0.tiles = dev[pos:pos + 4 * l0.n_tiles].view(l0.n_tiles, 4)


class MergedLayout:
    """Levels 0 and top of several consecutive chunk layouts viewed as one batch (rows of chunk i follow those of chunk
    i - 1, pair ids are renumbered): what the query decoder and the heads need when the backbone ran chunk by chunk.
    ``device_merge(levels, row_seq_out, seqinfo_out)`` (``CudaOps.merge_layout``) builds the device arrays from the chunk layouts
    already on the device; without it they are built on the host and uploaded (CPU tests, more than 16 chunks)."""

    def __init__(self, lays: Sequence[PackLayout], device, device_merge=None):
        self.n_levels = lays[0].n_levels
        self.B = sum(l.B for l in lays)
        self.lengths = np.concatenate([l.lengths for l in lays])
        self.levels: List[LevelLayout] = [None] * self.n_levels
        on_device = device_merge is not None and len(lays) <= 16 and torch.device(device).type == "cuda"
        host = []
        for lv in (0, self.n_levels - 1):
            row_base = np.cumsum([0] + [l.levels[lv].R for l in lays])
            off = np.concatenate([l.levels[lv].off + rb for l, rb in zip(lays, row_base)]).astype(np.int32)
            length = np.concatenate([l.levels[lv].len for l in lays]).astype(np.int32)
            haspad = np.concatenate([l.levels[lv].haspad for l in lays]).astype(np.int32)
            rows = int(row_base[-1])
            self.levels[lv] = LevelLayout(lv, off, length, haspad, rows)
            if on_device:
                lev = self.levels[lv]
                lev.row_seq = torch.empty(rows, dtype=torch.int32, device=device)
                lev.seqinfo = torch.empty(self.B, 4, dtype=torch.int32, device=device)
                device_merge([l.levels[lv] for l in lays], lev.row_seq, lev.seqinfo)
                continue
            pair_base = np.cumsum([0] + [l.B for l in lays])
            row_seq = np.concatenate([np.where(l._host[2 * lv] >= 0, l._host[2 * lv] + pb, -1) for l, pb in zip(lays, pair_base)])
            info = np.stack([off, length, haspad, np.zeros(self.B, np.int32)], 1)
            host += [row_seq.astype(np.int32), info.reshape(-1)]
        if on_device:
            return
        n_words = sum(h.size for h in host)
        if torch.device(device).type == "cuda":
            flat = torch.empty(n_words, dtype=torch.int32, pin_memory=True)
            np.concatenate(host, out=flat.numpy())
            dev = flat.to(device, non_blocking=True)
        else:
            dev = torch.from_numpy(np.concatenate(host))
        pos = 0
        for lv in (0, self.n_levels - 1):
            lev = self.levels[lv]
            lev.row_seq = dev[pos:pos + lev.R]
            pos += lev.R
            lev.seqinfo = dev[pos:pos + 4 * self.B].view(self.B, 4)
            pos += 4 * self.B
