"""CPU models of the bit tricks the CUDA kernels rely on (stringwars_b200/csrc/lev_bitpar*.cuh, common.cuh), checked
against the oracle: the column recurrence, the delta-bit hand-over between lanes, the register re-alignment of string
pieces and the shared-memory layout's bank behaviour. They pin the ALGEBRA on the CPU box; the kernels themselves are
checked bit-exact on the GPU by tests/test_gpu_parity.py."""
import itertools
import random

import numpy as np
import pytest

from oracle import pyref

M32 = 0xFFFFFFFF


def _eq_masks(pat, K):
    """Top-aligned pattern: symbol i at bit s + i, s = 32K - m (lev_bitpar.cuh)."""
    s = 32 * K - len(pat)
    eq = {}
    for i, c in enumerate(pat):
        p = s + i
        eq.setdefault(c, [0] * K)[p >> 5] |= 1 << (p & 31)
    return s, eq


def _first_column(s, K, first=0):
    Pv = []
    for w in range(K):
        lo = first + 32 * w
        Pv.append(0 if s >= lo + 32 else (M32 if s <= lo else (M32 << (s - lo)) & M32))
    return Pv, [0] * K


def column_step(eq, Pv, Mv, hp_in=1, hn_in=0, carry=False):
    """column_step_io: Hyyro's single-mask form over K words, ONE wide add, IMAD shift in word 0, funnel shifts above."""
    K = len(Pv)
    t, x = [], []
    for w in range(K):
        e = eq[w] | (hn_in if (carry and w == 0) else 0)
        t.append(e & Pv[w])
        x.append(e | Mv[w])
    total = sum(v << (32 * w) for w, v in enumerate(t)) + sum(v << (32 * w) for w, v in enumerate(Pv))
    ssum = [(total >> (32 * w)) & M32 for w in range(K)]
    ph_below = mh_below = 0
    for w in range(K):
        d0 = ((ssum[w] ^ Pv[w]) | x[w]) & M32
        ph = (Mv[w] | ~(d0 | Pv[w])) & M32
        mh = Pv[w] & d0
        if w == 0:
            phs, mhs = (ph * 2 + hp_in) & M32, (mh * 2 + (hn_in if carry else 0)) & M32
        else:
            phs, mhs = ((ph << 1) | (ph_below >> 31)) & M32, ((mh << 1) | (mh_below >> 31)) & M32
        ph_below, mh_below = ph, mh
        Pv[w] = (mhs | ~(d0 | phs)) & M32
        Mv[w] = phs & d0
    return ph_below, mh_below  # the top word's unshifted deltas: bit 31 leaves the block


def bitpar_distance(a, b, K):
    pat, txt = (a, b) if len(a) <= len(b) else (b, a)
    if not pat:
        return len(txt)
    s, eq = _eq_masks(pat, K)
    Pv, Mv = _first_column(s, K)
    for c in txt:
        column_step(eq.get(c, [0] * K), Pv, Mv)
    return len(txt) + sum(bin(v).count("1") for v in Pv) - sum(bin(v).count("1") for v in Mv)


def _rand_pair(rng, lo, hi, alphabet):
    a = bytes(rng.randrange(alphabet) for _ in range(rng.randint(lo, hi)))
    if rng.random() < 0.4:  # similar strings: long carry chains in the add
        b = bytearray(a)
        for _ in range(rng.randint(0, 4)):
            if b:
                b[rng.randrange(len(b))] = rng.randrange(alphabet)
        b = bytes(b) + bytes(rng.randrange(alphabet) for _ in range(rng.randint(0, 9)))
    else:
        b = bytes(rng.randrange(alphabet) for _ in range(rng.randint(lo, hi + 20)))
    return a, b


@pytest.mark.parametrize("K", [1, 2, 4, 8])
def test_column_step_matches_the_definition(K):
    rng = random.Random(K)
    for _ in range(150):
        a, b = _rand_pair(rng, 0, 32 * K, rng.choice([2, 4, 26]))
        if min(len(a), len(b)) > 32 * K:
            continue
        assert bitpar_distance(a, b, K) == pyref.levenshtein(a, b), (a, b)


def coop_distance(a, b, G, K):
    """lev_bitpar_coop.cuh: the column split over G lanes of K words, FOUR columns per wavefront step, the delta bits of
    a step packed most-significant-first into two nibbles (+1 bits in 3..0, -1 bits in 7..4)."""
    pat, txt = (a, b) if len(a) <= len(b) else (b, a)
    m, n = len(pat), len(txt)
    if m == 0:
        return n
    cap = 32 * G * K
    assert m <= cap
    s = cap - m
    lanes = []
    for lane in range(G):
        base = 32 * K * lane
        eq = {}
        for i, c in enumerate(pat):
            p = s + i - base
            if 0 <= p < 32 * K:
                eq.setdefault(c, [0] * K)[p >> 5] |= 1 << (p & 31)
        Pv, Mv = _first_column(s, K, base)
        lanes.append({"eq": eq, "Pv": Pv, "Mv": Mv, "live": base + 32 * K > s, "first": base <= s})
    groups = (n + 3) // 4
    send = [0] * G
    for t in range(groups + G - 1):
        recv = [0] + send[:-1]  # __shfl_up_sync(send, 1)
        new_send = list(send)
        for lane, L in enumerate(lanes):
            gg = t - lane
            if not L["live"] or not (0 <= gg < groups):
                continue
            inp = 0x0F if L["first"] else recv[lane]
            in_p, in_n = (inp << 28) & M32, ((inp >> 4) << 28) & M32
            out_p = out_n = 0
            for k in range(4):
                j = 4 * gg + k
                if j < n:
                    ph_top, mh_top = column_step(L["eq"].get(txt[j], [0] * K), L["Pv"], L["Mv"], in_p >> 31, in_n >> 31, carry=True)
                    out_p, out_n = ((out_p << 1) | (ph_top >> 31)) & M32, ((out_n << 1) | (mh_top >> 31)) & M32
                else:
                    out_p, out_n = (out_p << 1) & M32, (out_n << 1) & M32
                in_p, in_n = (in_p << 1) & M32, (in_n << 1) & M32
            new_send[lane] = (out_p | (out_n << 4)) & M32
        send = new_send
    delta = sum(bin(v).count("1") for L in lanes if L["live"] for v in L["Pv"]) - sum(
        bin(v).count("1") for L in lanes if L["live"] for v in L["Mv"])
    return n + delta


@pytest.mark.parametrize("G,K", [(4, 1), (8, 2), (4, 4)])
def test_lane_group_wavefront_with_four_column_steps(G, K):
    rng = random.Random(G * 10 + K)
    for _ in range(40):
        a, b = _rand_pair(rng, 1, 32 * G * K, rng.choice([2, 4, 26]))
        if min(len(a), len(b)) > 32 * G * K:
            continue
        assert coop_distance(a, b, G, K) == pyref.levenshtein(a, b), (a, b)


# ---- register re-alignment of 16-byte pieces (common.cuh: realign16, PieceStream)

def realign16(lo, hi, mis):
    """Two rounds of word selects (by 2 words, by 1 word) and four funnel shifts, on lists of four u32."""
    w = list(lo) + list(hi)
    by2, by1 = bool(mis & 8), bool(mis & 4)
    t = [w[i + 2] if by2 else w[i] for i in range(6)]
    u = [t[i + 1] if by1 else t[i] for i in range(5)]
    sh = (mis & 3) * 8
    return [(((u[i + 1] << 32) | u[i]) >> sh) & M32 for i in range(4)]


def _words(b16):
    return [int.from_bytes(b16[4 * i:4 * i + 4], "little") for i in range(4)]


def test_realign16_is_a_byte_shift_of_the_window():
    rng = np.random.default_rng(0)
    window = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    for mis in range(16):
        got = realign16(_words(window[:16]), _words(window[16:]), mis)
        assert got == _words(window[mis:mis + 16]), mis


def piece_stream(mem, start, length, lead):
    """PieceStream: yields (piece bytes, set of aligned piece numbers it loaded). `mem` is the whole address space; only
    aligned 16-byte pieces that hold at least one string byte may be read."""
    window = start - lead
    mis = window % 16
    grid = window - mis
    lo, hi = mis + lead, mis + lead + length

    def aligned(a):
        ok = 16 * a + 16 > lo and 16 * a < hi
        return (mem[grid + 16 * a: grid + 16 * a + 16] if ok else bytes(16)), ok

    i = 0
    while 16 * i < lead + length:
        p0, ok0 = aligned(i)
        loaded = {i} if ok0 else set()
        if mis:
            p1, ok1 = aligned(i + 1)
            loaded |= {i + 1} if ok1 else set()
            words = realign16(_words(p0), _words(p1), mis)
            piece = b"".join(x.to_bytes(4, "little") for x in words)
        else:
            piece = p0
        yield piece, {grid + 16 * a for a in loaded}
        i += 1


def test_piece_stream_reads_only_pieces_that_hold_string_bytes():
    rng = np.random.default_rng(1)
    mem = rng.integers(1, 256, 4096, dtype=np.uint8).tobytes()
    for start, length, lead in itertools.product((256, 257, 263, 271, 300), (1, 5, 16, 17, 40, 100), (0, 3, 11, 15)):
        out = bytearray()
        for piece, touched in piece_stream(mem, start, length, lead):
            out += piece
            for addr in touched:
                assert addr % 16 == 0 and addr + 16 > start and addr < start + length, (start, length, lead, addr)
        # window byte `lead + i` is string byte i, whatever the alignment
        assert bytes(out[lead:lead + length]) == mem[start:start + length], (start, length, lead)


# ---- shared-memory layout of the nibble tables (lev_bitpar.cuh: NibbleTables)

def _layout(K, T):
    parts = 1 if K <= 4 else K // 4
    sb = min(K, 4) * 4
    tpr = 128 // sb
    return parts, sb, tpr


def _base(K, T, tid):
    parts, sb, tpr = _layout(K, T)
    return (tid // tpr) * (parts * 2 * 2048) + (tid % tpr) * sb


@pytest.mark.parametrize("K,T", [(1, 64), (2, 64), (4, 64), (8, 64), (2, 128)])
def test_table_layout_is_a_partition_and_lookups_are_conflict_free(K, T):
    parts, sb, tpr = _layout(K, T)
    seen = set()
    for tid in range(T):
        for part in range(parts):
            for table in range(2):
                for entry in range(16):
                    a = _base(K, T, tid) + part * 4096 + table * 2048 + entry * 128
                    cells = set(range(a, a + sb))
                    assert not (cells & seen)
                    seen |= cells
    assert seen == set(range(T * K * 4 * 32))  # exactly NibbleTables::BYTES, no holes
    # a warp's lookup (every lane its own entry) is split by the hardware into wavefronts of 128 bytes worth of lanes:
    # within one wavefront the lanes' slots must fall into distinct banks whatever entries they ask for
    rng = random.Random(K)
    lanes_per_wavefront = 128 // sb
    for warp in range(T // 32):
        for _ in range(20):
            entries = [rng.randrange(16) for _ in range(32)]
            for wf in range(32 // lanes_per_wavefront):
                banks = []
                for lane in range(wf * lanes_per_wavefront, (wf + 1) * lanes_per_wavefront):
                    a = _base(K, T, warp * 32 + lane) + entries[lane] * 128
                    banks += [((a + 4 * i) // 4) % 32 for i in range(sb // 4)]
                assert len(set(banks)) == len(banks)


@pytest.mark.parametrize("K,T", [(2, 64), (4, 64), (8, 64)])
def test_build_rotation_keeps_a_uniform_batch_conflict_free(K, T):
    """Lanes whose slots share a bank start the build at different words: lane group g works on word (2q + g + p/2) % K."""
    parts, sb, tpr = _layout(K, T)
    sharers = 32 // tpr
    rng = random.Random(K)
    for q in range((K + 1) // 2):
        for p in range(4):
            banks = []
            for lane in range(32):
                g = (lane // tpr) % sharers
                word = (2 * q + g + p // 2) % K
                a = _base(K, T, lane) + (word >> 2) * 4096 + (word & 3) * 4 + rng.randrange(16) * 128
                banks.append((a // 4) % 32)
            assert len(set(banks)) == 32, (q, p)


# ---- packed s16x2 arithmetic of the affine / local wavefront (align_wavefront_packed_general.cuh)

def _s16(v):
    v &= 0xFFFF
    return v - 0x10000 if v & 0x8000 else v


def _halves(p):
    return _s16(p), _s16(p >> 16)


def _pack(lo, hi):
    return ((hi & 0xFFFF) << 16) | (lo & 0xFFFF)


def pk_both(v):
    return ((v & 0xFFFF) * 0x00010001) & M32


def viaddmax_s16x2(a, b, c):
    """VIADDMNMX.S16x2: per half max(a + b, c) with wrapping 16-bit adds (what the kernel relies on NOT happening)."""
    out = []
    for x, y, z in zip(_halves(a), _halves(b), _halves(c)):
        out.append(max(_s16(x + y), z))
    return _pack(*out)


def vmaxs2(a, b):
    return _pack(*[max(x, y) for x, y in zip(_halves(a), _halves(b))])


def vminu2(a, b):
    return _pack(*[min(x & 0xFFFF, y & 0xFFFF) for x, y in zip(_halves(a), _halves(b))])


def packed_general_scores(pair_a, pair_b, match, mismatch, first, nxt, local, affine):
    """Two pairs of the same shape, one in each half, the cell update of the kernel's `cells` lambda row by row."""
    (ra, ca), (rb, cb) = pair_a, pair_b
    m, n = len(ra), len(ca)
    assert (len(rb), len(cb)) == (m, n)
    NEG2, floor2 = 0x80008000, pk_both(-32000)
    first2, next2, match2, mismatch2 = pk_both(first), pk_both(nxt), pk_both(match), pk_both(mismatch)
    add = lambda a, b: viaddmax_s16x2(a, b, NEG2)
    Hprev = [pk_both(0 if (local or j == 0) else first + (j - 1) * nxt) for j in range(n + 1)]  # row 0
    Eprev = [floor2] * (n + 1)
    best = 0
    for i in range(1, m + 1):
        Hrow = [pk_both(0 if local else first + (i - 1) * nxt)] + [0] * n
        Erow = [floor2] * (n + 1)
        f = floor2
        rc = ra[i - 1] | (rb[i - 1] << 16)
        for j in range(1, n + 1):
            cc = ca[j - 1] | (cb[j - 1] << 16)
            ne = vminu2(rc ^ cc, 0x00010001)
            mask = (ne * 0xFFFF) & M32
            s = (mask & mismatch2) | (~mask & M32 & match2)
            if affine:
                e = viaddmax_s16x2(Eprev[j], next2, add(Hprev[j], first2))
                f = viaddmax_s16x2(f, next2, add(Hrow[j - 1], first2))
                h = vmaxs2(viaddmax_s16x2(Hprev[j - 1], s, e), f)
                Erow[j] = e
            else:
                h = vmaxs2(viaddmax_s16x2(Hprev[j - 1], s, add(Hrow[j - 1], next2)), add(Hprev[j], next2))
            if local:
                h = vmaxs2(h, 0)
                best = vmaxs2(best, h)
            Hrow[j] = h
        Hprev, Eprev = Hrow, Erow
    return _halves(best) if local else _halves(Hprev[n])


@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("first,nxt", [(-2, -2), (-5, -1), (-6, -1)])
def test_packed_s16x2_alignment_cells(local, first, nxt):
    rng = random.Random(first * 7 + nxt + local)
    affine = first != nxt
    sub = lambda x, y: 2 if x == y else -1
    for _ in range(25):
        m, n = rng.randint(1, 24), rng.randint(1, 40)
        mk = lambda k: bytes(rng.choice(b"ACGT") for _ in range(k))
        pa, pb = (mk(m), mk(n)), (mk(m), mk(n))
        got = packed_general_scores(pa, pb, 2, -1, first, nxt, local, affine)
        want = (pyref.align_score(pa[0], pa[1], sub, first, nxt, local), pyref.align_score(pb[0], pb[1], sub, first, nxt, local))
        assert tuple(got) == want, (pa, pb)


def test_packed_range_precondition_keeps_sums_inside_int16():
    """engine.cu routes a batch to the packed kernel only if Q = 3*first + (m+n)*min(next, mismatch) > -28000 and
    min(m,n)*match < 32000. Every H is at least the all-gaps score 2*first + (m+n)*min(next, mismatch) = Q - first, so
    with costs of int8 magnitude no add of the cell update can leave int16, not even from the gap tracks' floor."""
    floor = -32000
    for first, nxt, mismatch, match in itertools.product((-128, -12, -5, -2, 0), (-128, -3, -1, 0), (-128, -4, -1, 0), (0, 2, 127)):
        lo_cost = min(nxt, mismatch)
        if lo_cost == 0:
            continue
        total = (-28000 - 3 * first) // lo_cost  # the largest m + n the precondition admits (lo_cost < 0: floor division rounds towards it)
        while 3 * first + total * lo_cost <= -28000:
            total -= 1
        if total < 2:
            continue
        h_min = 2 * first + total * lo_cost
        h_max = 31999 if match == 0 else min(31999, (32000 // match) * match)
        for a, b in ((floor, nxt), (h_min, first), (h_min, nxt), (h_min, mismatch), (h_max, match), (floor, first)):
            assert -32768 <= a + b <= 32767, (first, nxt, mismatch, match, a, b)


# ------------------------------------------------------------------------------------------ round 2: lev_short.cuh, utf8.cuh

def _dp4a(a, b, c=0):
    return (sum(((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF) for i in range(4)) + c) & M32


def _eq_flags(word, c):
    """lev_short.cuh eq_flags: 0x80 in every byte of `word` equal to c — exact for all byte values (no borrow between bytes)."""
    t = word ^ (c * 0x01010101)
    s = ((t & 0x7F7F7F7F) + 0x7F7F7F7F) & M32
    return ~(s | t | 0x7F7F7F7F) & M32


def test_short_kernel_match_mask_in_registers():
    """The register kernel derives Eq(c) without tables: SWAR zero-byte detection, two IDP.4A that pack eight flags into a
    byte, the result positioned at bit 7 (seven free rows below the pattern). Exhaustive over all byte pairs for the
    detector, random patterns for the packing, then whole distances against the oracle's definition."""
    for a in range(256):
        word = a | (((a + 1) & 0xFF) << 8) | ((a ^ 0x80) << 16) | (0x01 << 24)  # neighbours that would trip a borrowing detector
        for c in range(256):
            want = sum(0x80 << (8 * i) for i in range(4) if ((word >> (8 * i)) & 0xFF) == c)
            assert _eq_flags(word, c) == want, (hex(word), c)
    rng = random.Random(3)

    def eq_mask7(pat_words, c):
        acc = 0
        for w in range(0, len(pat_words), 2):
            a = _dp4a(_eq_flags(pat_words[w], c), 0x08040201)
            if w + 1 < len(pat_words):
                a = _dp4a(_eq_flags(pat_words[w + 1], c), 0x80402010, a)
            acc = (acc + a * (256 ** (w // 2))) & M32
        return acc

    def short_distance(pat, txt, W):
        data = bytes(pat) + bytes(rng.randrange(256) for _ in range(4 * W - len(pat)))  # garbage above the pattern, as in a tape
        words = [int.from_bytes(data[4 * w: 4 * w + 4], "little") for w in range(W)]
        Pv, Mv = [0xFFFFFF80], [0]
        for c in txt:
            eq = eq_mask7(words, c)
            for i, p in enumerate(pat):  # bit 7 + i set iff pattern byte i == c
                assert ((eq >> (7 + i)) & 1) == (p == c)
            column_step([eq], Pv, Mv)
        rows = ((1 << len(pat)) - 1) << 7
        return len(txt) + bin(Pv[0] & rows).count("1") - bin(Mv[0] & rows).count("1")

    for _ in range(300):
        W = rng.choice([2, 4, 6])
        alpha = rng.choice([2, 4, 26, 256])
        pat = bytes(rng.randrange(alpha) for _ in range(rng.randrange(1, 4 * W + 1)))
        txt = bytes(rng.randrange(alpha) for _ in range(rng.randrange(0, 40)))
        assert short_distance(pat, txt, W) == pyref.levenshtein(pat, txt), (pat, txt, W)


def _lead_bits4(w):
    cont = (w >> 7) & ~(w >> 6) & 0x01010101
    lead = cont ^ 0x01010101
    return ((lead * 0x01020408) >> 24) & 0xF


def _rename_model(dict_str: bytes, other: bytes, ascii_identity: bool):
    """utf8.cuh utf8_rename_kernel, piece by piece as a lane sees it: 16 bytes + the 4 bytes after them, `ext` bits (a lead
    byte or the end of the string ends the symbol before it), little-endian packing of a lead byte and up to three
    continuation bytes, plain ASCII kept as it is, everything else numbered through a dictionary (the kernel numbers
    the newcomers of a 32-symbol step in lane order; any order gives the same distance). Returns the two renamed strings, or None when the dictionary overflows."""
    cap, id_base = (127, 128) if ascii_identity else (255, 0)
    numbers = {}  # the dictionary: packed symbol -> number, in the order the symbols are first met

    def lookup(key, insert):
        if key not in numbers:
            if not insert:
                return id_base
            if len(numbers) == cap:
                raise OverflowError
            numbers[key] = len(numbers) + 1
        return id_base + numbers[key]

    def rename(data, insert):
        out = bytearray()
        for at in range(0, len(data), 16):
            remaining = len(data) - at
            window = data[at: at + 20] + bytes(rng_garbage.randrange(256) for _ in range(20))  # past the end: somebody else's bytes
            words = [int.from_bytes(window[4 * q: 4 * q + 4], "little") for q in range(5)]
            if remaining <= 16:
                words[4] = 0  # no next piece: the kernel does not load one
            ext = sum(_lead_bits4(words[q]) << (4 * q) for q in range(5))
            if remaining < 20:
                ext |= (M32 << remaining) & M32
            valid = 0xFFFF if remaining >= 16 else (1 << remaining) - 1
            leads = ext & valid
            high = (words[0] | words[1] | words[2] | words[3]) & 0x80808080
            if ascii_identity and high == 0 and (ext & 0x1FFFF) == 0x1FFFF:
                out += window[:bin(leads).count("1")]  # the fast path stages 16 bytes; `total` keeps the symbols only
                continue
            for j in range(16):
                if not (leads >> j) & 1:
                    continue
                four = int.from_bytes(window[j: j + 4], "little")
                z = ((ext >> (j + 1)) & 7) | 8
                conts = (z & -z).bit_length() - 1
                if ascii_identity and conts == 0 and not four & 0x80:
                    out.append(four & 0x7F)
                else:
                    out.append(lookup(four & (M32 >> (24 - 8 * conts)), insert))
        return bytes(out)

    try:
        return rename(dict_str, True), rename(other, False)
    except OverflowError:
        return None


rng_garbage = random.Random(77)


def _tape(strings):
    data = np.frombuffer(b"".join(strings) + b"\0", dtype=np.uint8)[:-1].copy()
    offs = np.zeros(len(strings) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in strings])
    return data, offs


def test_utf8_symbol_split_and_rename_model():
    """utf8.cuh: lead bits of a word (a byte starts a symbol iff it is not 10xxxxxx), and the per-pair rename — ASCII kept,
    other symbols numbered by dictionary slot, symbols the shorter string does not hold collapsed to one id — preserves the
    code-point distance, on valid text and on arbitrary bytes (the oracle's lenient split)."""
    from oracle import oracle

    rng = random.Random(9)
    for _ in range(2000):
        bs = bytes(rng.choice([0x41, 0x7F, 0x80, 0xBF, 0xC3, 0xE4, 0xF0, 0x00, 0xFF]) for _ in range(4))
        assert _lead_bits4(int.from_bytes(bs, "little")) == sum(1 << i for i in range(4) if (bs[i] & 0xC0) != 0x80), bs

    pool = "aбвжzß日本語😀é x"
    nasty = [0x41, 0x00, 0x7F, 0x80, 0xBF, 0xC3, 0xE4, 0xF0, 0xFF, 0x20]
    for trial in range(400):
        if trial % 2:
            a = "".join(rng.choice(pool) for _ in range(rng.randrange(0, 40))).encode()
            b = "".join(rng.choice(pool) for _ in range(rng.randrange(0, 40))).encode()
            want = pyref.levenshtein_utf8(a, b)
        else:  # not UTF-8 at all: stray continuation bytes, overlong runs of them, lone leads, strings that end inside a symbol
            a = bytes(rng.choice(nasty) for _ in range(rng.randrange(0, 50)))
            b = bytes(rng.choice(nasty) for _ in range(rng.randrange(0, 50)))
            want = int(oracle.levenshtein_pairs(*_tape([a]), *_tape([b]), oracle.UTF8)[0])
        for ascii_identity in (True, False):
            d, o = (a, b) if len(a) <= len(b) else (b, a)
            ren = _rename_model(d, o, ascii_identity)
            assert ren is not None
            assert pyref.levenshtein(ren[0], ren[1]) == want, (a, b, ascii_identity)
    # capacity: 127 distinct multi-byte symbols fit the first tier, 128 do not; the second tier takes 255 symbols of any kind
    wide = "".join(chr(0x4E00 + i) for i in range(300))
    assert _rename_model(wide[:127].encode(), wide.encode(), True) is not None
    assert _rename_model(wide[:128].encode(), wide.encode(), True) is None
    assert _rename_model((wide[:200] + "abc").encode(), wide.encode(), False) is not None
    assert _rename_model(wide[:256].encode(), wide.encode(), False) is None


def test_utf8_batch_dictionary_model():
    """utf8.cuh, cross products: one dictionary per batch. Plain ASCII keeps its value, every other symbol (big-endian packing
    of a lead byte and up to three continuation bytes) is numbered from 129 in whatever order the survey meets it; the
    translated byte tapes have the code-point distances of the originals for EVERY pair of the cross product, and a batch with
    more than 127 distinct multi-byte symbols does not fit."""
    from oracle import oracle

    def symbols(data):  # symbol_at() over the lead bytes of a string
        out, i = [], 0
        while i < len(data):
            if (data[i] & 0xC0) == 0x80:
                i += 1
                continue
            sym, k = data[i], 1
            while k < 4 and i + k < len(data) and (data[i + k] & 0xC0) == 0x80:
                sym = (sym << 8) | data[i + k]
                k += 1
            out.append(sym)
            i += 1
        return out

    def translate(strings, order_rng):
        numbers = {}
        seen = [s for x in strings for s in symbols(x) if s >= 0x80]
        order_rng.shuffle(seen)  # atomics: any insertion order
        for s in seen:
            if s not in numbers:
                numbers[s] = len(numbers) + 1
        if len(numbers) > 127:
            return None
        return [bytes(s if s < 0x80 else 128 + numbers[s] for s in symbols(x)) for x in strings]

    rng = random.Random(21)
    pool = "abcxyz éèßøñ жукпривет"
    nasty = [0x41, 0x00, 0x80, 0xBF, 0xC3, 0xF0]
    for trial in range(6):
        if trial % 2:
            strings = ["".join(rng.choice(pool) for _ in range(rng.randrange(0, 14))).encode() for _ in range(16)]
        else:
            strings = [bytes(rng.choice(nasty) for _ in range(rng.randrange(0, 14))) for _ in range(16)]
        ren = translate(strings, rng)
        assert ren is not None
        q, c = strings[:6], strings[6:]
        want = oracle.levenshtein_cross(*_tape(q), *_tape(c), oracle.UTF8)
        for i in range(6):
            for j in range(10):
                assert pyref.levenshtein(ren[i], ren[6 + j]) == want[i, j], (q[i], c[j])
    wide = ["".join(chr(0x4E00 + i) for i in range(k, k + 16)).encode() for k in range(0, 128, 16)]
    assert translate(wide[:7] + ["日本語".encode()], rng) is not None  # 112 + 3 distinct
    assert translate(wide + ["x".encode()], rng) is None               # 128 distinct ideographs: one too many
