"""Host-mirror support library (librlr_hostmirror.so: rlr_lexical_*, host_mirror/lexical.cpp) against the
pure-Python restatement of the reference's LexicalIndex (oracle/lexical.py).  No GPU needed: the index is
host code in the reference too (src/rag_engine.rs:2083-2247).  It is NOT part of the product library."""
import ctypes as C
import random

import numpy as np

from oracle import lexical as olex


def _lib(rlr):
    return rlr.load_hostmirror()


def _tok(lib, text):
    b = text.encode("utf-8")
    out = C.create_string_buffer(len(b) * 2 + 16)
    n, nt = C.c_size_t(0), C.c_uint32(0)
    assert lib.rlr_tokenize(b, len(b), out, len(out), C.byref(n), C.byref(nt)) == 0
    s = out.raw[:n.value].decode("utf-8")
    toks = s.split("\n") if n.value else []
    assert len(toks) == nt.value
    return toks


def _score(lib, lx, query, limit):
    b = query.encode("utf-8")
    cap = 4096
    keys, scores, n = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_uint32(0)
    rc = lib.rlr_lexical_score(lx, b, len(b), limit, keys.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p), cap, C.byref(n))
    assert rc == 0
    return [(int(k), np.float32(s)) for k, s in zip(keys[:n.value], scores[:n.value])]


def test_tokenizer_matches_restatement(rlr):
    lib = _lib(rlr)
    cases = [
        "The quick brown fox jumps over the lazy dog.",
        "", "a bc def", "ab,cd;efg-HIJ_klm", "x86_64 ARMv8 3.14159 2024-01-15", "URLs: https://example.com/path?q=1",
        "naïve café ÉCOLE Ünïcödé straße", "ΑΘΗΝΑ αθήνα Ελλάδα", "МОСКВА москва Привет, мир!", "日本語のテキスト、漢字。",
        "tab\tseparated\nnew\r\nlines", "émigré—dash…ellipsis “quoted” ‘single’", "№5 ½ cup x² 10µm", "aé", "éé", "ab",
        "Ǆ Ā ā Ĳ ĳ Ŀ ŀ Ÿ Ž ž", "full-width：ＡＢＣ！？",
        # punctuation of other scripts must split tokens (Arabic comma / question mark, Hebrew maqaf,
        # Armenian full stop, Devanagari danda), and their letters must be case-folded like Rust does
        "مرحبا،بالعالم؟نعم", "שלום־עולם", "Բարև։Աշխարհ ԲԱՐԵՎ", "नमस्ते।दुनिया हिन्दी", "ღმერთი ᲦᲛᲔᲠᲗᲘ Ⴀⴀ",
        "ΟΔΥΣΣΕΥΣ ΟΔΟΣ Σ ΑΣ.ΣΑ ΣΟΦΙΑ", "İSTANBUL ISPARTA ıspanak", "Ǳǲǳ ǅ Ꙁꙁ Ԁԁ Ⱥ Ɐ", "𠀀𠀁𠀂 𐐀𐐨 𝐀𝐁𝐂",   # supplementary planes
        "e\u0301cole E\u0301COLE a\u0345b", "ⅠⅡⅢ ①②③ ㈠ ½¾",
    ]
    for text in cases:
        assert _tok(lib, text) == olex.tokenize(text), text


def test_unicode_tables_agree_with_their_sources_on_every_code_point(rlr):
    """Every code point up to U+10FFFF: is_alphanumeric == Alphabetic | Nd | Nl | No (the `regex` module's database)
    and the lowercase mapping == CPython's str.lower().  Guards the generated tables (tools/gen_unicode_tables.py)
    and their binary searches; VERDICT r1 / ADVICE r1: the hand-written ranges mis-classified whole scripts."""
    import regex
    lib = _lib(rlr)
    n = 0x110000
    alnum = np.zeros(n, np.uint8)
    lower = np.zeros(3 * n, np.uint32)
    assert lib.rlr_hostmirror_unicode_dump(alnum.ctypes.data_as(C.c_void_p), lower.ctypes.data_as(C.c_void_p), n) == 0
    pat = regex.compile(r"[\p{Alphabetic}\p{Nd}\p{Nl}\p{No}]")
    lower = lower.reshape(n, 3)
    bad = []
    for cp in range(n):
        if 0xD800 <= cp <= 0xDFFF:
            assert alnum[cp] == 0
            continue
        ch = chr(cp)
        want_alnum = pat.match(ch) is not None
        want_lower = [ord(x) for x in ch.lower()]
        got_lower = [int(x) for x in lower[cp] if x] or [0]
        if bool(alnum[cp]) != want_alnum or got_lower != want_lower:
            bad.append(hex(cp))
    assert not bad, bad[:20]
    # the cases the hand-written tables got wrong
    for cp in (0x060C, 0x061F, 0x05BE, 0x0589, 0x0964):
        assert alnum[cp] == 0, hex(cp)
    for cp in (0x20000, 0x2A6DF, 0x10400, 0x1D400, 0x0561, 0x10D0):
        assert alnum[cp] == 1, hex(cp)
    assert lower[0x0130].tolist() == [0x69, 0x307, 0] and lower[0x0531].tolist() == [0x0561, 0, 0]


def _corpus(seed, n_docs, vocab):
    rng = random.Random(seed)
    docs = []
    for _ in range(n_docs):
        n = rng.randint(0, 60)
        docs.append(" ".join(rng.choice(vocab) for _ in range(n)))
    return docs


VOCAB = ("retrieval augmented generation embedding vector cosine similarity rust tokio axum server pdf chunk sentence "
         "overlap index search query rerank lexical bm25 the and of to in a is it on at by an be GPU kernel HBM "
         "bandwidth tensor Memory memory MEMORY naïve café straße Ελλάδα москва 2024 42 x86").split()


def test_bm25_scores_bit_identical_to_restatement(rlr):
    lib = _lib(rlr)
    for seed in (1, 2, 3):
        docs = _corpus(seed, 300, VOCAB)
        lx = C.c_void_p()
        assert lib.rlr_lexical_create(C.byref(lx)) == 0
        ref = olex.LexicalIndex()
        for i, d in enumerate(docs):
            b = d.encode("utf-8")
            assert lib.rlr_lexical_add_chunk(lx, 1000 + i, b, len(b)) == 0
            ref.add_chunk(1000 + i, d)
        td, tl, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        assert lib.rlr_lexical_stats(lx, C.byref(td), C.byref(tl), C.byref(nt)) == 0
        assert (td.value, tl.value, nt.value) == (ref.total_docs, ref.total_length, len(ref.term_postings))
        for query, limit in (("memory bandwidth of the GPU kernel", 25), ("rust tokio axum", 0), ("zzz unknown", 10),
                             ("a an it", 10), ("Memory MEMORY memory", 1500), ("café naïve москва", 7), ("", 5)):
            got, want = _score(lib, lx, query, limit), ref.score(query, limit)
            assert [k for k, _ in got] == [k for k, _ in want], (seed, query)
            assert all(np.float32(a).tobytes() == np.float32(b).tobytes() for (_, a), (_, b) in zip(got, want)), (seed, query)
        lib.rlr_lexical_destroy(lx)


def test_add_replace_remove_bookkeeping(rlr):
    """add_chunk on an existing key replaces it (:2107-2109); remove_chunk undoes it (:2140-2167)."""
    lib = _lib(rlr)
    lx = C.c_void_p()
    lib.rlr_lexical_create(C.byref(lx))
    ref = olex.LexicalIndex()
    docs = _corpus(9, 120, VOCAB)
    rng = random.Random(4)
    for step in range(600):
        key = rng.randrange(80)
        if rng.random() < 0.6:
            d = rng.choice(docs)
            b = d.encode("utf-8")
            lib.rlr_lexical_add_chunk(lx, key, b, len(b))
            ref.add_chunk(key, d)
        else:
            lib.rlr_lexical_remove_chunk(lx, key)
            ref.remove_chunk(key)
        if step % 50 == 49:
            td, tl, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
            lib.rlr_lexical_stats(lx, C.byref(td), C.byref(tl), C.byref(nt))
            assert (td.value, tl.value, nt.value) == (ref.total_docs, ref.total_length, len(ref.term_postings))
            got, want = _score(lib, lx, "embedding search memory kernel", 0), ref.score("embedding search memory kernel", 0)
            assert [(k, np.float32(s).tobytes()) for k, s in got] == [(k, np.float32(s).tobytes()) for k, s in want]
            for k in range(80):
                c = C.c_int(-1)
                lib.rlr_lexical_contains(lx, k, C.byref(c))
                assert bool(c.value) == (k in ref.doc_terms)
    lib.rlr_lexical_destroy(lx)


def test_empty_and_argument_errors(rlr):
    lib = _lib(rlr)
    lx = C.c_void_p()
    lib.rlr_lexical_create(C.byref(lx))
    assert _score(lib, lx, "anything at all", 10) == []          # total_docs == 0 -> empty (:2170-2172)
    lib.rlr_lexical_add_chunk(lx, 1, b"a b c", 5)                 # no token survives the >= 3 bytes filter -> not indexed
    c = C.c_int(-1)
    lib.rlr_lexical_contains(lx, 1, C.byref(c))
    assert c.value == 0
    n = C.c_uint32(0)
    assert lib.rlr_lexical_score(None, b"x", 1, 1, None, None, 0, C.byref(n)) == rlr.RLR_ERR_INVALID_ARG
    lib.rlr_lexical_add_chunk(lx, 2, b"alpha beta gamma", 16)
    assert lib.rlr_lexical_score(lx, b"alpha", 5, 0, None, None, 0, C.byref(n)) == rlr.RLR_ERR_UNSUPPORTED   # no room
    lib.rlr_lexical_destroy(lx)
