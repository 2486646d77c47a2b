"""Deterministic synthetic inputs for the BASELINE.json configs (SURVEY.md 8d / BASELINE.md 2).

PRNG: splitmix64 seeds a bank of xoshiro256** lanes (vectorised in numpy, so the byte stream does not depend on the
numpy version).  seed = 0x5A170000 + config * 0x100 + buffer index.

Classes of the "Silesia-mix" (buffer i uses class i mod 8):
  0 English-like text        1 XML-like tagged text       2 source-code / log-like lines   3 x86-like binary
  4 fixed-width DB records   5 16-bit smooth signal+noise 6 order-2 Markov low-entropy     7 high entropy + motifs
"""
import numpy as np

U64 = np.uint64
_M = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return x, z ^ (z >> 31)


class Rng:
    """LANES independent xoshiro256** generators advanced in lock step."""

    LANES = 4096

    def __init__(self, seed):
        s = np.empty((4, self.LANES), dtype=U64)
        x = seed & _M
        for lane in range(self.LANES):
            for k in range(4):
                x, v = _splitmix64(x)
                s[k, lane] = v
        self.s = s
        self._buf = np.empty(0, dtype=U64)

    def _step(self):
        s = self.s
        with np.errstate(over="ignore"):
            r = s[1] * U64(5)
            r = ((r << U64(7)) | (r >> U64(57))) * U64(9)
            t = s[1] << U64(17)
            s[2] ^= s[0]
            s[3] ^= s[1]
            s[1] ^= s[2]
            s[0] ^= s[3]
            s[2] ^= t
            s[3] = (s[3] << U64(45)) | (s[3] >> U64(19))
        return r

    def u64(self, n):
        parts = [self._buf]
        have = self._buf.size
        while have < n:
            r = self._step()
            parts.append(r)
            have += r.size
        allv = np.concatenate(parts)
        self._buf = allv[n:]
        return allv[:n]

    def uniform(self, n):
        return (self.u64(n) >> U64(11)).astype(np.float64) * (1.0 / (1 << 53))

    def integers(self, lo, hi, n):
        return (lo + (self.uniform(n) * (hi - lo)).astype(np.int64)).astype(np.int64)

    def bytes(self, n):
        return self.u64((n + 7) // 8).view(np.uint8)[:n].copy()

    def choice_p(self, cdf, n):
        return np.searchsorted(cdf, self.uniform(n), side="right").astype(np.int64)


def _table(strings):
    """list of bytes -> (matrix [k, maxlen] uint8, lens [k])"""
    lens = np.array([len(s) for s in strings], dtype=np.int64)
    m = np.zeros((len(strings), max(1, int(lens.max()))), dtype=np.uint8)
    for i, s in enumerate(strings):
        m[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    return m, lens


def _ragged(mat, lens, idx):
    """concatenate mat[idx[j], :lens[idx[j]]] for all j"""
    l = lens[idx]
    total = int(l.sum())
    starts = np.cumsum(l) - l
    tok = np.repeat(np.arange(idx.size), l)
    off = np.arange(total) - starts[tok]
    return mat[idx[tok], off]


_VOCAB = None


def _vocab():
    global _VOCAB
    if _VOCAB is None:
        r = Rng(0x5A17FFFF)
        lens = r.integers(3, 11, 2048)
        letters_cdf = np.cumsum(np.array([8.2, 1.5, 2.8, 4.3, 12.7, 2.2, 2.0, 6.1, 7.0, 0.15, 0.8, 4.0, 2.4, 6.7, 7.5, 1.9,
                                          0.1, 6.0, 6.3, 9.1, 2.8, 1.0, 2.4, 0.15, 2.0, 0.07]))
        letters_cdf /= letters_cdf[-1]
        words = []
        for L in lens:
            words.append(bytes((97 + r.choice_p(letters_cdf, int(L))).astype(np.uint8)))
        w = 1.0 / np.arange(1, 2049) ** 1.1
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        _VOCAB = (words, cdf)
    return _VOCAB


def gen_text(n, seed):
    """class 0: Zipf(1.1) words of a fixed 2048-word vocabulary, '. ' every 8-20 words, newline about every 80 chars."""
    words, cdf = _vocab()
    r = Rng(seed)
    seps = [b" ", b". ", b"\n", b".\n"]
    toks = [w + s for s in seps for w in words]  # index = sep * 2048 + word
    mat, lens = _table(toks)
    nw = n // 5 + 64
    widx = r.choice_p(cdf, nw)
    u = r.uniform(nw)
    sep = np.zeros(nw, dtype=np.int64)
    sep[u < 1.0 / 14] = 1       # sentence end
    u2 = r.uniform(nw)
    nl = u2 < 1.0 / 13          # newline roughly every 80 characters
    sep = sep + 2 * nl.astype(np.int64)
    out = _ragged(mat, lens, sep * 2048 + widx)
    while out.size < n:
        out = np.concatenate([out, out])
    return out[:n].copy()


def gen_xml(n, seed):
    """class 1: tagged text with repeated tag and attribute names."""
    words, cdf = _vocab()
    r = Rng(seed)
    tags = [b"item", b"entry", b"record", b"node", b"field", b"value", b"name", b"title", b"para", b"link", b"ref", b"section"]
    attrs = [b"id", b"type", b"class", b"lang", b"href", b"ts", b"ver"]
    toks = []
    for t in tags:
        for a in attrs:
            toks.append(b"  <" + t + b" " + a + b'="')
    n_open = len(toks)
    for t in tags:
        toks.append(b"</" + t + b">\n")
    n_close = len(tags)
    base_num = len(toks)
    for v in range(1000):
        toks.append(b"%d" % v)
    base_word = len(toks)
    for w in words[:1024]:
        toks.append(w + b" ")
    q_close = len(toks)
    toks.append(b'">')
    mat, lens = _table(toks)
    nrec = n // 40 + 16
    t_idx = r.integers(0, len(tags), nrec)
    a_idx = r.integers(0, len(attrs), nrec)
    num = r.integers(0, 1000, nrec)
    k = 6
    wsel = np.minimum(r.choice_p(cdf, nrec * k), 1023).reshape(nrec, k)
    seq = np.empty((nrec, 4 + k), dtype=np.int64)
    seq[:, 0] = t_idx * len(attrs) + a_idx
    seq[:, 1] = base_num + num
    seq[:, 2] = q_close
    seq[:, 3:3 + k] = base_word + wsel
    seq[:, 3 + k] = n_open + t_idx
    out = _ragged(mat, lens, seq.reshape(-1))
    while out.size < n:
        out = np.concatenate([out, out])
    return out[:n].copy()


_LEVELS = [b"INFO", b"WARN", b"ERROR", b"DEBUG", b"TRACE"]


def _log_tables():
    words, _ = _vocab()
    hosts = [b"host%02d" % i for i in range(64)]
    svcs = [b"svc-" + words[100 + i] for i in range(32)]
    tmpl = []
    r = Rng(0x5A17FFFE)
    for i in range(256):
        k = int(r.integers(3, 9, 1)[0])
        ws = r.integers(0, 512, k)
        tmpl.append(b" ".join(words[int(j)] for j in ws))
    return hosts, svcs, tmpl


def gen_log(n, seed, t0=1577836800000):
    """class 2 / config C4: '<ISO-8601 ts> <level> <host> <svc> req=<hex16> <template> [numeric fields]' lines with a
    monotonically increasing millisecond timestamp (+1..2000 ms per line)."""
    hosts, svcs, tmpl = _log_tables()
    r = Rng(seed)
    nl = n // 90 + 16
    dt = r.integers(1, 2001, nl)
    ts = t0 + np.cumsum(dt)
    # timestamp text: 2020-01-01T00:00:00.000Z style from epoch milliseconds (all dates stay within 2020 for < 3e10 ms)
    ms = ts % 1000
    sec = (ts // 1000)
    days = (sec // 86400 - 18262).astype(np.int64)  # days since 2020-01-01
    sod = sec % 86400
    mdays = np.array([31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31] * 4)
    cum = np.concatenate([[0], np.cumsum(mdays)])
    yday = days % 366
    mon = np.searchsorted(cum[1:13], yday, side="right")
    dom = yday - cum[mon] + 1
    year = 2020 + days // 366

    def digits(v, w):
        v = v.astype(np.int64)
        cols = [(v // (10 ** (w - 1 - i))) % 10 + 48 for i in range(w)]
        return np.stack(cols, axis=1).astype(np.uint8)

    nlines = nl
    parts = [digits(year, 4), np.full((nlines, 1), ord("-"), np.uint8), digits(mon + 1, 2), np.full((nlines, 1), ord("-"), np.uint8),
             digits(dom, 2), np.full((nlines, 1), ord("T"), np.uint8), digits(sod // 3600, 2), np.full((nlines, 1), ord(":"), np.uint8),
             digits((sod // 60) % 60, 2), np.full((nlines, 1), ord(":"), np.uint8), digits(sod % 60, 2),
             np.full((nlines, 1), ord("."), np.uint8), digits(ms, 3), np.full((nlines, 1), ord("Z"), np.uint8),
             np.full((nlines, 1), ord(" "), np.uint8)]
    tsmat = np.concatenate(parts, axis=1)  # [nl, 25]
    hexd = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    req = hexd[(r.u64(nl)[:, None] >> (U64(4) * np.arange(16, dtype=U64)[None, :])).astype(np.int64) & 15]
    toks = []
    base_lvl = 0
    toks += [l + b" " for l in _LEVELS]
    base_host = len(toks)
    toks += [h + b" " for h in hosts]
    base_svc = len(toks)
    toks += [s + b" req=" for s in svcs]
    base_t = len(toks)
    toks += [b" " + t for t in tmpl]
    base_num = len(toks)
    toks += [b" %d" % v for v in range(2000)]
    empty = len(toks)
    toks.append(b"")
    nlt = len(toks)
    toks.append(b"\n")
    mat, lens = _table(toks)
    lvl = r.choice_p(np.array([0.7, 0.8, 0.85, 0.97, 1.0]), nl)
    host = r.integers(0, 64, nl)
    svc = r.integers(0, 32, nl)
    tm = r.integers(0, 256, nl)
    nf = r.integers(0, 4, nl)
    nums = r.integers(0, 2000, nl * 3).reshape(nl, 3)
    # head tokens (level host svc) | req hex | tail tokens (template, up to 3 numbers, newline)
    head = _ragged(mat, lens, np.stack([base_lvl + lvl, base_host + host, base_svc + svc], axis=1).reshape(-1))
    head_len = lens[base_lvl + lvl] + lens[base_host + host] + lens[base_svc + svc]
    tail_idx = np.stack([base_t + tm,
                         np.where(nf > 0, base_num + nums[:, 0], empty),
                         np.where(nf > 1, base_num + nums[:, 1], empty),
                         np.where(nf > 2, base_num + nums[:, 2], empty),
                         np.full(nl, nlt)], axis=1)
    tail = _ragged(mat, lens, tail_idx.reshape(-1))
    tail_len = lens[tail_idx].sum(axis=1)
    line_len = 25 + head_len + 16 + tail_len
    total = int(line_len.sum())
    out = np.empty(total, dtype=np.uint8)
    start = np.cumsum(line_len) - line_len

    def scatter(src_flat, seg_len, seg_start):
        tok = np.repeat(np.arange(seg_len.size), seg_len)
        off = np.arange(int(seg_len.sum())) - np.repeat(np.cumsum(seg_len) - seg_len, seg_len)
        out[seg_start[tok] + off] = src_flat

    scatter(tsmat.reshape(-1), np.full(nl, 25), start)
    scatter(head, head_len, start + 25)
    scatter(req.reshape(-1), np.full(nl, 16), start + 25 + head_len)
    scatter(tail, tail_len, start + 25 + head_len + 16)
    while out.size < n:
        out = np.concatenate([out, out])
    return out[:n].copy()


def gen_x86(n, seed):
    """class 3: opcode bytes from a skewed table, ModRM bytes, and 4-byte little-endian addresses with small deltas."""
    r = Rng(seed)
    ni = n // 4 + 16
    ops = np.array([0x8B, 0x89, 0xE8, 0x83, 0xFF, 0x8D, 0x85, 0x74, 0x75, 0x0F, 0xC7, 0x50, 0x51, 0x53, 0x55, 0x56, 0x57, 0x5D, 0xC3,
                    0x33, 0x3B, 0xEB, 0x6A, 0x68, 0xA1, 0xB8], dtype=np.uint8)
    w = 1.0 / np.arange(1, ops.size + 1)
    cdf = np.cumsum(w) / w.sum()
    op = ops[r.choice_p(cdf, ni)]
    modrm = r.bytes(ni) & np.uint8(0xC7) | np.uint8(0x05)
    has_imm = r.uniform(ni) < 0.35
    addr = (0x00401000 + np.cumsum(r.integers(-64, 256, ni))).astype(np.int64) & 0xFFFFFFFF
    rec = np.zeros((ni, 6), dtype=np.uint8)
    rec[:, 0] = op
    rec[:, 1] = modrm
    for k in range(4):
        rec[:, 2 + k] = (addr >> (8 * k)) & 0xFF
    lens = np.where(has_imm, 6, np.where(r.uniform(ni) < 0.5, 2, 1)).astype(np.int64)
    out = _ragged(rec, lens, np.arange(ni))
    while out.size < n:
        out = np.concatenate([out, out])
    return out[:n].copy()


def gen_db(n, seed):
    """class 4: 64-byte fixed-width records with sorted decimal keys and small-vocabulary fields."""
    words, cdf = _vocab()
    r = Rng(seed)
    nr = n // 64 + 2
    key = 10000000 + np.cumsum(r.integers(1, 40, nr))
    rec = np.full((nr, 64), ord(" "), dtype=np.uint8)
    for i in range(8):
        rec[:, i] = (key // (10 ** (7 - i))) % 10 + 48
    rec[:, 8] = ord("|")
    wm, wl = _table([w.ljust(10)[:10] for w in words[:256]])
    f1 = np.minimum(r.choice_p(cdf, nr), 255)
    f2 = np.minimum(r.choice_p(cdf, nr), 255)
    rec[:, 9:19] = wm[f1]
    rec[:, 19] = ord("|")
    rec[:, 20:30] = wm[f2]
    rec[:, 30] = ord("|")
    amt = r.integers(0, 1000000, nr)
    for i in range(7):
        rec[:, 31 + i] = (amt // (10 ** (6 - i))) % 10 + 48
    rec[:, 38] = ord("|")
    flag = r.integers(0, 4, nr)
    rec[:, 39] = np.frombuffer(b"YNAX", dtype=np.uint8)[flag]
    rec[:, 63] = ord("\n")
    return rec.reshape(-1)[:n].copy()


def gen_signal(n, seed):
    """class 5: 16-bit little-endian smooth signal (random walk of a slowly varying slope) plus noise."""
    r = Rng(seed)
    ns = n // 2 + 2
    slope = np.cumsum(r.integers(-3, 4, ns))
    slope = np.clip(slope - int(slope.mean()), -200, 200).astype(np.int64)
    sig = np.cumsum(slope) // 16 + r.integers(-6, 7, ns)
    v = (sig & 0xFFFF).astype(np.uint16)
    return v.view(np.uint8)[:n].copy()


def gen_markov(n, seed):
    """class 6: order-2 Markov bytes over a 32-symbol alphabet with 4 skewed successors per context."""
    r = Rng(seed)
    chains = 1024
    steps = n // chains + 2
    tr = Rng(0x5A17FFFD)
    succ = (tr.integers(0, 32, 32 * 32 * 4).reshape(32, 32, 4) + 32).astype(np.uint8)
    a = np.full(chains, 32, dtype=np.int64)
    b = np.full(chains, 33, dtype=np.int64)
    out = np.empty((steps, chains), dtype=np.uint8)
    cdf = np.array([0.6, 0.85, 0.95, 1.0])
    ch = r.choice_p(cdf, steps * chains).reshape(steps, chains)
    for t in range(steps):
        c = succ[a - 32, b - 32, ch[t]]
        out[t] = c
        a, b = b, c.astype(np.int64)
    return out.T.reshape(-1)[:n].copy()


def gen_entropy(n, seed):
    """class 7: 90 % random bytes + 10 % repeated 64-byte motifs."""
    r = Rng(seed)
    out = r.bytes(n)
    motifs = Rng(0x5A17FFFC).bytes(16 * 64).reshape(16, 64)
    nm = max(1, n // 640)
    pos = r.integers(0, max(1, n - 64), nm)
    which = r.integers(0, 16, nm)
    for p, w in zip(pos.tolist(), which.tolist()):
        out[p:p + 64] = motifs[w][:max(0, min(64, n - p))]
    return out


_CLASSES = [gen_text, gen_xml, gen_log, gen_x86, gen_db, gen_signal, gen_markov, gen_entropy]
CLASS_NAMES = ["text", "xml", "log", "x86", "db", "signal", "markov", "entropy"]


def seed_for(config, index):
    return 0x5A170000 + config * 0x100 + index


def silesia_mix(index, n, config=3):
    """buffer `index` of the Silesia-mix configs (C3/C5): class index mod 8."""
    return _CLASSES[index % 8](n, seed_for(config, index))


def text_buffer(index, n, config=2):
    """C1/C2 text buffers."""
    return gen_text(n, seed_for(config, index))


def _log_piece(args):
    k, n, config = args
    # piece k of a long stream: its own seed, its timestamps behind the pieces in front of it (+1..2000 ms per ~90-byte line)
    return gen_log(n, seed_for(config, k), t0=1577836800000 + k * (LOG_PIECE // 90 + 16) * 1001)


LOG_PIECE = 32 << 20


def log_stream(n, config=4, workers=None):
    """C4: one long log stream.  Up to 64 MiB it is one generator run; longer streams are pieces of 32 MiB (each its own seed,
    timestamps still increasing across the pieces) generated by a few processes -- the single run takes minutes per GiB."""
    if n <= (64 << 20):
        return gen_log(n, seed_for(config, 0))
    import os
    from concurrent.futures import ProcessPoolExecutor
    pieces = [(k, min(LOG_PIECE, n - k * LOG_PIECE), config) for k in range((n + LOG_PIECE - 1) // LOG_PIECE)]
    if workers is None:
        try:
            workers = len(os.sched_getaffinity(0))
        except AttributeError:
            workers = os.cpu_count() or 1
        workers = max(1, min(workers, 32, len(pieces)))
    out = np.empty(n, dtype=np.uint8)
    with ProcessPoolExecutor(max_workers=workers) as ex:
        for (k, m, _), piece in zip(pieces, ex.map(_log_piece, pieces)):
            out[k * LOG_PIECE:k * LOG_PIECE + m] = piece
    return out
