"""Host-side planning logic of the C layer (safconv_host.c): FFT size, partition count, MAC tiling and
the split-K work distribution / gather tables.  Pure CPU: calls safconv_debug_plan, no CUDA."""
import ctypes as C

import numpy as np
import pytest


class Plan(C.Structure):   # mirrors scdev_plan in csrc/safconv_dev.h
    _fields_ = [("kind", C.c_int), ("hop", C.c_int), ("len", C.c_int), ("nIn", C.c_int), ("nOutLocal", C.c_int),
                ("N", C.c_int), ("M", C.c_int), ("logM", C.c_int), ("P", C.c_int), ("fftThreads", C.c_int),
                ("nKT", C.c_int), ("nOT", C.c_int), ("OTsz", C.c_int), ("SNI", C.c_int), ("SPU", C.c_int),
                ("R", C.c_int), ("WGo", C.c_int), ("WGk", C.c_int), ("totalStages", C.c_longlong),
                ("macGrid", C.c_int), ("nGroups", C.c_int), ("nSlots", C.c_int), ("RS", C.c_int), ("maxBatch", C.c_int),
                ("macHints", C.c_int),
                ("macSmemBytes", C.c_int), ("macStages", C.c_int), ("macStageBytes", C.c_int),
                ("nIRs", C.c_int)]


def plan(saf, hop, L, nIn, nOut, sms=148, kind=0):
    lib = saf.lib()
    assert lib.safconv_debug_plan_size() == C.sizeof(Plan), "tests/test_plan.py Plan mirror is out of date"
    cap = 1 << 16
    pl = Plan()
    cta = (C.c_int * cap)()
    gs = (C.c_int * cap)()
    gl = (C.c_int * cap)()
    rc = lib.safconv_debug_plan(kind, hop, L, nIn, nOut, sms, C.byref(pl), cta, gs, gl, cap)
    assert rc == 0
    return pl, np.array(cta[:pl.macGrid + 1]), np.array(gs[:pl.nGroups + 1]), np.array(gl[:pl.nSlots])


CASES = [
    (256, 1024, 4, 2), (128, 512, 25, 2), (2048, 512, 32, 40), (1024, 96000, 64, 64), (1024, 96000, 64, 8),
    (96, 250, 3, 5), (100, 333, 2, 3), (64, 64, 1, 1), (512, 4096, 7, 65), (1024, 8192, 121, 64), (8192, 100, 2, 130),
    (1, 5, 2, 2), (17, 40, 3, 3),
]


@pytest.mark.parametrize("hop,L,nIn,nOut", CASES)
def test_fft_and_partition_plan(saf, hop, L, nIn, nOut):
    pl, *_ = plan(saf, hop, L, nIn, nOut)
    assert pl.N >= 2 * hop and pl.N >= 64 and (pl.N & (pl.N - 1)) == 0
    assert pl.N < 4 * hop or pl.N == 64                   # smallest such power of two
    assert pl.M == pl.N // 2 and (1 << pl.logM) == pl.M
    assert pl.P == int(np.ceil(np.float32(L) / np.float32(hop)))   # reference .c:102
    assert pl.fftThreads % 32 == 0 and 32 <= pl.fftThreads <= 256


@pytest.mark.parametrize("hop,L,nIn,nOut", CASES)
@pytest.mark.parametrize("sms", [148, 7])
def test_mac_tiling_and_split_tables(saf, hop, L, nIn, nOut, sms):
    pl, ctaBase, grpStart, grpList = plan(saf, hop, L, nIn, nOut, sms)
    # tiling covers all outputs / inputs
    assert pl.nOT * pl.OTsz >= nOut and pl.OTsz <= 64
    assert 1 <= pl.R <= 8 and pl.WGo in (1, 2, 4, 8) and pl.WGo * pl.R >= pl.OTsz and pl.WGo * pl.WGk <= 8 and pl.WGk >= 1
    assert pl.SNI * pl.SPU >= nIn and (pl.SPU - 1) * pl.SNI < nIn
    assert pl.SNI * pl.OTsz * 256 <= pl.macStageBytes or pl.SNI == 1
    assert pl.nKT * 32 == pl.M and pl.nGroups == pl.nOT * pl.nKT
    assert pl.totalStages == pl.nGroups * pl.P * pl.SPU
    assert 1 <= pl.macGrid <= min(sms, pl.totalStages)
    assert pl.macSmemBytes <= 227 * 1024
    # emulate the kernel's work split and check it against the gather tables
    T, G, spg = pl.totalStages, pl.macGrid, pl.P * pl.SPU
    produced = {}                                   # slot id -> (group, stages covered)
    covered = np.zeros(pl.nGroups, np.int64)
    for c in range(G):
        s0, s1 = T * c // G, T * (c + 1) // G
        assert s1 > s0
        it = s0
        while it < s1:
            g = it // spg
            end = min(s1, (g + 1) * spg)
            slot = ctaBase[c] + (g - s0 // spg)
            assert slot not in produced
            produced[slot] = g
            covered[g] += end - it
            it = end
    assert len(produced) == pl.nSlots == ctaBase[G]
    assert np.all(covered == spg)                   # every stage of every group streamed exactly once
    assert grpStart[0] == 0 and grpStart[-1] == pl.nSlots
    for g in range(pl.nGroups):
        slots = grpList[grpStart[g]:grpStart[g + 1]]
        assert len(slots) >= 1
        assert all(produced[s] == g for s in slots)
    # the slots of a group are consecutive: K3 sums slots grpStart[g] .. grpStart[g+1]-1 in ascending order
    assert grpList.tolist() == list(range(pl.nSlots))
    assert pl.maxBatch >= 1 and pl.RS == pl.P + pl.maxBatch


@pytest.mark.parametrize("hop,L,nIn,nOut", [(1024, 96000, 64, 64), (256, 3000, 6, 10), (128, 2000, 3, 70), (512, 1024, 33, 17),
                                           (64, 128, 2, 3)])
@pytest.mark.parametrize("which", ["tail", "head"])
def test_lookahead_pass_tables(saf, hop, L, nIn, nOut, which):
    """Split tables of the look-ahead passes (tail = partitions 1..P-1, head = partition 0): emulate the MAC kernel's
    walk over a pass -- stage -> (group, partition, stage in unit) with the filter pointer skipping the partitions
    outside the pass -- and check that every (group, partition, stage) of the pass is streamed exactly once, that the
    filter offsets are the ones of the full layout, and that the gather table lists each group's slots consecutively."""
    lib = saf.lib()
    pl, *_ = plan(saf, hop, L, nIn, nOut)
    P, SPU, nG = pl.P, pl.SPU, pl.nGroups
    pLo, nP = (1, P - 1) if which == "tail" else (0, 1)
    cap = 1 << 16
    cta = (C.c_int * cap)(); gs = (C.c_int * cap)(); gl = (C.c_int * cap)()
    T = C.c_longlong(); G = C.c_int()
    slots = lib.safconv_debug_pass_tables(hop, L, nIn, nOut, 148, pLo, nP, C.byref(T), C.byref(G), cta, gs, gl, cap)
    assert slots > 0
    T, G = T.value, G.value
    assert T == nG * nP * SPU and 1 <= G <= min(148, T)
    ctaBase = np.array(cta[:G + 1]); grpStart = np.array(gs[:nG + 1]); grpList = np.array(gl[:slots])
    unit_elems = nIn * pl.OTsz * 32                      # float2 elements of one (group, partition) unit of H
    seen = set()
    slot_group = {}
    for c in range(G):
        s0, s1 = T * c // G, T * (c + 1) // G
        unit0 = s0 // SPU
        sidx, grp0 = s0 - unit0 * SPU, unit0 // nP
        p = pLo + (unit0 - grp0 * nP)
        g = grp0
        srcH = (grp0 * P + p) * unit_elems + sidx * pl.SNI * pl.OTsz * 32       # kernel: srcH0
        slot = ctaBase[c]
        for _ in range(s1 - s0):
            ni0 = sidx * pl.SNI
            cnt = min(pl.SNI, nIn - ni0)
            assert srcH == (g * P + p) * unit_elems + ni0 * pl.OTsz * 32         # offset in the full [group][P][nIn][OTsz][32] layout
            assert (g, p, sidx) not in seen
            seen.add((g, p, sidx))
            slot_group.setdefault(slot, g)
            assert slot_group[slot] == g
            srcH += cnt * pl.OTsz * 32
            sidx += 1
            if sidx == SPU:
                sidx = 0
                p += 1
                if p == pLo + nP:
                    p = pLo
                    srcH += (P - nP) * unit_elems        # skip the partitions outside the pass
                    g += 1
                    slot += 1                            # kernel: dst advances at the end of a group
    assert len(seen) == nG * nP * SPU
    assert grpStart[0] == 0 and grpStart[-1] == slots and grpList.tolist() == list(range(slots))
    for g in range(nG):
        for sl in range(grpStart[g], grpStart[g + 1]):
            assert slot_group.get(sl, g) == g
