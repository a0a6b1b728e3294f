"""Host logic of the one-sided peer-memory transport (v3d_b200.viewshard.PeerTransport) on CPU.

The transport's kernels only copy bytes and raise / poll flag words; everything that can go wrong on the host side is
ADDRESS ARITHMETIC: arena offsets that must agree on all ranks although the frame blocks are uneven, the position of a
halo frame inside the NEIGHBOUR's pad buffer (which depends on the neighbour's block length), the slot of a rank in every
other rank's gather buffer, the rotation of the big buffers, the per-site flag words and the site cursor that lets a
captured graph replay the same addresses.  Here all ranks of a group live in ONE process: every rank gets a real
`PeerTransport` whose arena chunks are host buffers and whose C-ABI calls are replaced by a recorder that executes the
copies (`v3d_peer_put` -> memmove + flag stores; `v3d_peer_allreduce_f64` -> slot writes, summed once every rank has
called; `v3d_peer_wait` -> recorded and verified afterwards: every flag a rank waits on must have been raised, with the
current epoch, by exactly the rank it expects).  The exchanges then run through the public `ViewShard` / `CfgSplit`
methods the modules call, twice (the second pass must reuse the sites of the first), and the received data is compared
with what an all-gather / halo exchange / all-reduce must deliver.
"""
import ctypes as C
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from v3d_b200 import viewshard  # noqa: E402
from v3d_b200.viewshard import CfgSplit, PeerTransport, ViewShard, partition_frames  # noqa: E402


class _Group:
    """All ranks of one process group inside this process: shared arena registry + the fake C ABI."""

    def __init__(self, world: int):
        self.world = world
        self.chunks = []            # chunk index -> [host tensor per rank]
        self.ranks = []
        self.pending_reduce = {}    # site id -> list of (rank, stats ptr, n, scale, slot bases, flag bases)
        self.waits = []             # (rank, [flag addresses], epoch at the time of the wait)
        self.epoch = [0] * world

    def make(self, rank: int) -> PeerTransport:
        t = PeerTransport.__new__(PeerTransport)
        t.lib, t.check = _FakeLib(self, rank), (lambda rc, what: None)
        t.group, t.dev = None, torch.device("cpu")
        t.rank, t.world = rank, self.world
        t.chunks, t.cur, t.off = [], -1, 0
        t.ctrl = torch.zeros(64, dtype=torch.int32)
        t.scopes, t.scope, t.pos, t.rot, t.nsites = {}, "", 0, {}, 0
        grp = self

        def new_chunk(need, t=t):
            nbytes = max(1 << 20, (need + 0xFFFFF) & ~0xFFFFF)
            idx = len(t.chunks)
            if idx == len(grp.chunks):          # first rank to get here allocates the chunk for every rank
                grp.chunks.append([torch.zeros(nbytes, dtype=torch.uint8) for _ in range(grp.world)])
            bufs = grp.chunks[idx]
            assert bufs[0].numel() == nbytes, "ranks disagree on a chunk size: arena layouts diverged"
            t.chunks.append((nbytes, [b.data_ptr() for b in bufs], bufs[t.rank]))
            t.cur, t.off = idx, 0

        t._new_chunk = new_chunk
        self.ranks.append(t)
        return t

    def finish_reduces(self):
        for site, calls in self.pending_reduce.items():
            assert len(calls) == self.world, f"all-reduce site {site}: {len(calls)} of {self.world} ranks called"
            for rank, ptr, n, scale, slots, flags in calls:
                mine = (C.c_double * (n * self.world)).from_address(slots[rank])
                out = (C.c_double * n).from_address(ptr)
                for i in range(n):
                    out[i] = sum(mine[r * n + i] for r in range(self.world)) * scale
        self.pending_reduce.clear()

    def check_waits(self):
        for rank, flags, epoch in self.waits:
            for f in flags:
                assert C.c_uint32.from_address(f).value == epoch, f"rank {rank} waits on a flag nobody raised (epoch {epoch})"
        self.waits.clear()


class _FakeLib:
    def __init__(self, grp: _Group, rank: int):
        self.g, self.rank = grp, rank

    def v3d_peer_epoch_bump(self, epoch_ptr, stream):
        self.g.epoch[self.rank] += 1
        return 0

    def v3d_peer_put(self, n, src, dst, nb, m, fl, epoch, counter, stream):
        for i in range(n):
            C.memmove(dst[i], src[i], nb[i])
        for i in range(m):
            C.c_uint32.from_address(fl[i]).value = self.g.epoch[self.rank]
        return 0

    def v3d_peer_wait(self, n, fl, epoch, status, site, stream):
        self.g.waits.append((self.rank, [fl[i] for i in range(n)], self.g.epoch[self.rank]))
        return 0

    def v3d_peer_allreduce_f64(self, stats, n, scale, world, rank, slot, flag, epoch, status, site, stream):
        src = (C.c_double * n).from_address(stats)
        for r in range(world):
            dst = (C.c_double * (n * world)).from_address(slot[r])
            for i in range(n):
                dst[rank * n + i] = src[i]
            C.c_uint32.from_address(flag[r] + 4 * rank).value = self.g.epoch[self.rank]
        self.g.waits.append((self.rank, [flag[rank] + 4 * r for r in range(world)], self.g.epoch[self.rank]))
        self.g.pending_reduce.setdefault(site, []).append((rank, stats, n, scale, [slot[r] for r in range(world)], None))
        return 0


@pytest.fixture(autouse=True)
def _no_cuda_stream(monkeypatch):
    monkeypatch.setattr(viewshard, "_stream_of", lambda dev: 0)


def _shards(T: int, world: int):
    grp = _Group(world)
    blocks = partition_frames(T, world)
    shards = []
    for r in range(world):
        vs = ViewShard(T, r, world, None, blocks)
        vs.peer = grp.make(r)
        shards.append(vs)
    return grp, shards


@pytest.mark.parametrize("T,world", [(7, 3), (18, 4), (5, 2), (18, 8)])
def test_halo_gather_allreduce_addresses_with_uneven_blocks(T, world):
    grp, shards = _shards(T, world)
    nb, hw, c = 2, 4, 8
    video = torch.arange(nb * T * hw * c, dtype=torch.float32).reshape(nb, T, hw, c).to(torch.bfloat16)
    for pass_ in range(2):                    # the second pass must land on the same sites / buffers
        for vs in shards:
            vs.begin("unet")
        # (1) statistics all-reduce: local sums -> global sums * tl / T
        stats = [torch.full((nb, 32, 2), float(vs.rank + 1), dtype=torch.float64) for vs in shards]
        for vs, st in zip(shards, stats):
            vs.allreduce_stats_(st)
        grp.finish_reduces()
        total = sum(r + 1 for r in range(world))
        for vs, st in zip(shards, stats):
            assert torch.allclose(st, torch.full_like(st, total * vs.tl / T))
        # (2) halo exchange on two shapes (two rotating buffers each)
        for width in (c, 2 * c):
            vid = video if width == c else torch.cat([video, -video], dim=-1)
            pads = []
            for vs in shards:
                pad = vs.new_pad((nb, vs.tl + 2, hw, width), torch.bfloat16, "cpu")
                pad[:, 1:vs.tl + 1] = vid[:, vs.t0:vs.t0 + vs.tl]
                pad[:, 0] = 99.0
                pad[:, vs.tl + 1] = 99.0
                pads.append(pad)
            for vs, pad in zip(shards, pads):
                vs.exchange_halos(pad)
            for vs, pad in zip(shards, pads):
                left = vid[:, vs.t0 - 1] if vs.t0 > 0 else torch.zeros_like(vid[:, 0])
                right = vid[:, vs.t0 + vs.tl] if vs.t0 + vs.tl < T else torch.zeros_like(vid[:, 0])
                assert torch.equal(pad[:, 0], left) and torch.equal(pad[:, vs.tl + 1], right), (vs.rank, width)
                assert torch.equal(pad[:, 1:vs.tl + 1], vid[:, vs.t0:vs.t0 + vs.tl])
        # (3) K|V gather: every rank packs its rows into its slot; afterwards every buffer holds every block
        bufs = []
        for vs in shards:
            buf, mine = vs.kv_slots(nb, hw, 2 * c, torch.bfloat16, "cpu")
            rows = nb * vs.tl * hw
            mine[:rows] = torch.cat([video, video + 1], -1)[:, vs.t0:vs.t0 + vs.tl].reshape(rows, 2 * c)
            bufs.append((buf, mine, rows))
        for vs, (buf, mine, rows) in zip(shards, bufs):
            assert vs.kv_fused(buf) == vs.peer.kv_destinations(buf)            # destinations of the fused GEMM scatter
            assert vs.kv_fused(buf)[vs.rank] == mine.data_ptr()
            vs.gather_rows(mine, buf, rows)
        kv = torch.cat([video, video + 1], -1)
        for vs, (buf, _, _) in zip(shards, bufs):
            row, bstride = vs.kv_table(nb, hw)
            for f in range(T):
                for b in range(nb):
                    got = buf[row[f] + b * bstride[f]: row[f] + b * bstride[f] + hw]
                    assert torch.equal(got, kv[b, f]), (vs.rank, f, b)
        grp.check_waits()
        if pass_ == 0:
            nsites = [vs.peer.nsites for vs in shards]
            layout = [[(st["kind"], st.get("buf"), st.get("flags"), st.get("slots")) for st in vs.peer.scopes["unet"]]
                      for vs in shards]
    assert [vs.peer.nsites for vs in shards] == nsites, "the second pass created new sites"
    assert all(l == layout[0] for l in layout), "arena layouts differ between ranks"
    # the two pads of one shape alternate between two buffers; a third use would return to the first
    pad_sites = [st for st in shards[0].peer.scopes["unet"] if st["kind"] == "pad"]
    assert len({st["buf"] for st in pad_sites}) == 2 and pad_sites[0]["buf"] != pad_sites[1]["buf"]


def test_site_order_change_is_detected():
    grp, shards = _shards(6, 2)
    vs = shards[0]
    vs.begin("unet")
    vs.new_pad((1, vs.tl + 2, 4, 8), torch.bfloat16, "cpu")
    vs.begin("unet")
    with pytest.raises(RuntimeError, match="launch schedule must be static"):
        vs.kv_slots(1, 4, 16, torch.bfloat16, "cpu")


def test_cfg_pair_gather_alternates_two_sites():
    grp = _Group(2)
    pair = [CfgSplit(r, None, 0, grp.make(r)) for r in range(2)]
    for step in range(4):
        halves = [torch.full((3, 4, 2, 2), float(10 * step + r)) for r in range(2)]
        outs = []
        for cs, h in zip(pair, halves):
            cs.exchanges += 1
            parity = cs.exchanges & 1
            cs.peer.begin(f"cfg{parity}")
            outs.append(cs.peer.gather_halves(h, parity))
        for out in outs:
            assert torch.equal(out[:3], halves[0]) and torch.equal(out[3:], halves[1])
        grp.check_waits()
    assert pair[0].peer.nsites == 2 and set(pair[0].peer.scopes) == {"cfg0", "cfg1"}
