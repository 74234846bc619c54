"""Deterministic synthetic corpora for the benchmark configs of SURVEY.md section 8(d).

There is no network and no dataset in the image, so every workload is generated from a fixed
seed.  numpy only (no torch, no CUDA): this module is shared by bench.py, the tests and the
reference arm so that both arms compress exactly the same bytes.
"""
from __future__ import annotations

import numpy as np

SEED_ZIPF_TEXT = 0xB2130001
SEED_SOURCE = 0xB2130002
SEED_MIXED = 0xB2130003
SEED_LOG = 0xB2130004

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])


def _flatten(tokens: np.ndarray, tab_bytes: np.ndarray, tab_off: np.ndarray, tab_len: np.ndarray) -> np.ndarray:
    """Concatenate table entries tab[tokens[k]] into one byte array (vectorised gather)."""
    out = []
    step = 1 << 20  # chunked so the index temporaries stay cache-sized; int32 indices
    tab_len32 = tab_len.astype(np.int32)
    tab_off32 = tab_off.astype(np.int32)
    for a in range(0, len(tokens), step):
        tk = tokens[a:a + step]
        lens = tab_len32[tk]
        total = int(lens.sum())
        dst_start = np.cumsum(lens, dtype=np.int32) - lens
        idx = np.repeat(tab_off32[tk] - dst_start, lens)
        idx += np.arange(total, dtype=np.int32)
        out.append(tab_bytes[idx])
    return np.concatenate(out) if len(out) != 1 else out[0]


def _make_table(strings):
    lens = np.array([len(s) for s in strings], dtype=np.int64)
    off = np.cumsum(lens) - lens
    return np.frombuffer(b"".join(strings), dtype=np.uint8), off, lens


def _vocab(rng, nwords, lo, hi, alphabet=_LETTERS, weights=_LETTER_W):
    lens = rng.integers(lo, hi + 1, size=nwords)
    total = int(lens.sum())
    p = weights / weights.sum()
    flat = alphabet[rng.choice(len(alphabet), size=total, p=p)]
    off = np.cumsum(lens) - lens
    return flat, off.astype(np.int64), lens.astype(np.int64)


def _zipf_ids(rng, nvocab, s, count):
    w = 1.0 / np.power(np.arange(1, nvocab + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    # 2^16-entry guide table turns the search into one lookup plus a short local search
    u = rng.random(count)
    ids = np.searchsorted(cdf, u, side="right")
    return np.minimum(ids, nvocab - 1).astype(np.int64)


def zipf_text(nbytes: int, seed: int = SEED_ZIPF_TEXT) -> np.ndarray:
    """Config 2: 'enwik8-style' Zipf(1.1) text over a 65 536-word vocabulary with wiki-ish markup."""
    rng = np.random.default_rng(seed)
    vb, vo, vl = _vocab(rng, 65536, 2, 12)
    seps = [b" ", b", ", b". ", b"\n", b" [[", b"]] ", b" <ref>", b"</ref> "]
    sb, so, sl = _make_table(seps)
    tab_bytes = np.concatenate([vb, sb])
    tab_off = np.concatenate([vo, so + len(vb)])
    tab_len = np.concatenate([vl, sl])
    out = []
    have = 0
    while have < nbytes:
        nw = max(1024, int((nbytes - have) / 6.5) + 1024)
        words = _zipf_ids(rng, 65536, 1.1, nw)
        r = rng.random(nw)
        sep = np.where(r < 0.85, 0, np.where(r < 0.91, 1, np.where(r < 0.96, 2, 3)))
        m = rng.random(nw) < (1.0 / 200.0)  # sprinkle markup roughly every 200 words
        sep = np.where(m, 4 + rng.integers(0, 4, size=nw), sep)
        toks = np.empty(2 * nw, dtype=np.int64)
        toks[0::2] = words
        toks[1::2] = 65536 + sep
        chunk = _flatten(toks, tab_bytes, tab_off, tab_len)
        out.append(chunk)
        have += len(chunk)
    return np.concatenate(out)[:nbytes].copy()


def source_corpus(nbytes: int, seed: int = SEED_SOURCE) -> np.ndarray:
    """Config 3: C-like source files assembled from function templates; 30 % near-duplicate files."""
    rng = np.random.default_rng(seed)
    ident_b, ident_o, ident_l = _vocab(rng, 20000, 3, 14)
    kw = [b"int ", b"static ", b"void ", b"return ", b"if (", b") {\n", b"}\n", b"    ", b" = ", b";\n", b"(", b")",
          b", ", b" + ", b" - ", b" * ", b"->", b"/* ", b" */\n", b"for (", b"while (", b" < ", b"++", b"0", b"1",
          b"NULL", b"sizeof(", b"u32 ", b"const ", b"\n"]
    kb, ko, kl = _make_table(kw)
    tab_bytes = np.concatenate([ident_b, kb])
    tab_off = np.concatenate([ident_o, ko + len(ident_b)])
    tab_len = np.concatenate([ident_l, kl])
    K = 20000

    def template(r):
        n_stmt = int(r.integers(4, 40))
        ids = _zipf_ids(r, 20000, 1.05, 6 * n_stmt + 8)
        t = [K + 1, K + 2, ids[0], K + 10, K + 0, ids[1], K + 11, K + 5]
        p = 2
        for _ in range(n_stmt):
            kind = int(r.integers(0, 5))
            t += [K + 7]
            if kind == 0:
                t += [ids[p], K + 8, ids[p + 1], K + 13 + int(r.integers(0, 3)), ids[p + 2], K + 9]
            elif kind == 1:
                t += [K + 4, ids[p], K + 21, ids[p + 1], K + 5, K + 7, K + 7, K + 3, ids[p + 2], K + 9, K + 7, K + 6]
            elif kind == 2:
                t += [K + 17, ids[p], ids[p + 1], ids[p + 2], K + 18]
            elif kind == 3:
                t += [ids[p], K + 10, ids[p + 1], K + 12, ids[p + 2], K + 11, K + 9]
            else:
                t += [K + 19, ids[p], K + 21, ids[p + 1], K + 5, K + 7, K + 7, ids[p + 2], K + 22, K + 9, K + 7, K + 6]
            p += 3
        t += [K + 6, K + 29]
        return np.array(t, dtype=np.int64)

    templates = [template(rng) for _ in range(2048)]
    files = []
    out = []
    have = 0
    while have < nbytes:
        if files and rng.random() < 0.30:
            base = files[int(rng.integers(0, len(files)))].copy()
            nedit = max(1, int(len(base) * 0.015))
            pos = rng.integers(0, len(base), size=nedit)
            base[pos] = rng.integers(97, 123, size=nedit).astype(np.uint8)
            f = base
        else:
            nfun = int(rng.integers(8, 64))
            pick = _zipf_ids(rng, 2048, 0.8, nfun)
            toks = np.concatenate([templates[int(i)] for i in pick])
            f = _flatten(toks, tab_bytes, tab_off, tab_len)
        if len(files) < 4096:
            files.append(f)
        out.append(f)
        have += len(f)
    return np.concatenate(out)[:nbytes].copy()


def mixed(nbytes: int, seed: int = SEED_MIXED, segment: int = 64 << 20) -> np.ndarray:
    """Config 4: alternating segments of text, u32 delta arrays, 16-bit PCM-like walk, ~10 % random."""
    rng = np.random.default_rng(seed)
    out = []
    have = 0
    k = 0
    while have < nbytes:
        seg = min(segment, nbytes - have)
        if rng.random() < 0.10:
            b = rng.integers(0, 256, size=seg, dtype=np.uint8)
        elif k % 3 == 0:
            b = zipf_text(seg, seed=seed + 17 * k + 1)
        elif k % 3 == 1:
            d = rng.integers(0, 9, size=seg // 4 + 1).astype(np.uint32)
            b = np.cumsum(d, dtype=np.uint32).view(np.uint8)[:seg]
        else:
            w = np.cumsum(rng.integers(-40, 41, size=seg // 2 + 1)).astype(np.int16)
            b = w.view(np.uint8)[:seg]
        out.append(np.ascontiguousarray(b))
        have += seg
        k += 1
    return np.concatenate(out)[:nbytes].copy()


def log_stream(nbytes: int, seed: int = SEED_LOG) -> np.ndarray:
    """Config 5: timestamp + level + component + templated message, monotone timestamps."""
    rng = np.random.default_rng(seed)
    comp_b, comp_o, comp_l = _vocab(rng, 64, 4, 10)
    word_b, word_o, word_l = _vocab(rng, 4096, 2, 9)
    levels = [b" INFO ", b" WARN ", b" DEBUG ", b" ERROR "]
    lines = []
    have = 0
    t = 1_700_000_000_000
    templates = []
    for _ in range(512):
        nw = int(rng.integers(3, 12))
        ws = _zipf_ids(rng, 4096, 1.0, nw)
        templates.append(b" ".join(bytes(word_b[word_o[w]:word_o[w] + word_l[w]]) for w in ws))
    while have < nbytes:
        batch = 20000
        dt = rng.integers(0, 50, size=batch)
        lv = rng.choice(4, size=batch, p=[0.7, 0.1, 0.17, 0.03])
        cp = rng.integers(0, 64, size=batch)
        tp = _zipf_ids(rng, 512, 0.9, batch)
        hx = rng.integers(0, 1 << 32, size=batch)
        ip = rng.integers(0, 256, size=(batch, 4))
        for i in range(batch):
            t += int(dt[i])
            s, ms = divmod(t, 1000)
            d, rem = divmod(s, 86400)
            hh, rem = divmod(rem, 3600)
            mm, ss = divmod(rem, 60)
            c = int(cp[i])
            line = b"2026-%02d-%02dT%02d:%02d:%02d.%03dZ" % (1 + (d // 28) % 12, 1 + d % 28, hh, mm, ss, ms)
            line += levels[int(lv[i])] + bytes(comp_b[comp_o[c]:comp_o[c] + comp_l[c]]) + b": "
            line += templates[int(tp[i])] + b" id=%08x src=%d.%d.%d.%d\n" % (int(hx[i]), *[int(x) for x in ip[i]])
            lines.append(line)
            have += len(line)
            if have >= nbytes:
                break
    return np.frombuffer(b"".join(lines), dtype=np.uint8)[:nbytes].copy()


# ---------------------------------------------------------------- small edge-case inputs for parity tests
def edge_cases(seed: int = 12345):
    """(name, bytes) pairs covering the cases the reference's fuzzers and seeds exercise."""
    rng = np.random.default_rng(seed)
    cases = []

    def add(name, arr):
        cases.append((name, np.ascontiguousarray(arr, dtype=np.uint8).tobytes()))

    add("empty", np.zeros(0, np.uint8))
    add("one", np.array([7], np.uint8))
    add("raw63", rng.integers(0, 256, 63))
    add("coded64", rng.integers(0, 256, 64))
    add("coded65_text", np.frombuffer(b"the quick brown fox jumps over the lazy dog, twice over the lazy dog!!", np.uint8)[:65])
    add("below_lzp_min_71", rng.integers(0, 4, 71))
    add("lzp_min_72", rng.integers(0, 4, 72))
    add("zeros_4k", np.zeros(4096, np.uint8))
    add("zeros_100k", np.zeros(100_000, np.uint8))
    add("ff_run_70000", np.full(70_000, 255, np.uint8))
    add("random_10k", rng.integers(0, 256, 10_000))
    add("random_300k", rng.integers(0, 256, 300_000))
    add("two_symbols_50k", rng.integers(0, 2, 50_000))
    add("period2_60k", np.tile(np.array([97, 98], np.uint8), 30_000))
    add("period7_70k", np.tile(np.frombuffer(b"abcdefg", np.uint8), 10_000))
    add("period300_90k", np.tile(rng.integers(0, 256, 300), 300))
    fib_a, fib_b = b"a", b"ab"
    while len(fib_b) < 120_000:
        fib_a, fib_b = fib_b, fib_b + fib_a
    add("fibonacci_120k", np.frombuffer(fib_b[:120_000], np.uint8))
    runs = np.repeat(rng.integers(0, 256, 3000), rng.integers(1, 700, 3000))
    add("long_runs", runs[:200_000])
    runs255 = np.repeat(np.array([5, 9, 5, 9, 5], np.uint8), [255, 256, 510, 511, 1])
    add("run_boundaries_255", np.concatenate([runs255, rng.integers(0, 256, 100)]))
    esc = rng.integers(0, 256, 80_000)
    esc[rng.integers(0, 80_000, 6000)] = 0xF2
    add("escape_heavy", esc)
    rep = rng.integers(0, 256, 5000)
    add("repeat_block_5000x20", np.tile(rep, 20))
    near = np.tile(rng.integers(97, 123, 4000), 25)
    near[rng.integers(0, len(near), 300)] = 32
    add("near_repeats", near)
    add("zipf_200k", zipf_text(200_000, seed=99))
    add("log_150k", log_stream(150_000, seed=98))
    lz = np.concatenate([np.tile(rng.integers(0, 256, 41), 50), rng.integers(0, 256, 500),
                         np.tile(rng.integers(0, 256, 43), 50)])
    add("matches_len_40_43", lz)
    return cases
