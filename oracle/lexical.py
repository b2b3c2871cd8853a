"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's LexicalIndex (BM25) and tokenizer.

Follows /root/reference/src/rag_engine.rs:2083-2247 line by line, in pure Python with numpy
float32 scalars (small cases only).  Parity status: the reference has NO test that pins BM25
scores or tokenizer output, so this restatement is pinned by source text
only.  Where the reference is nondeterministic (HashSet/HashMap iteration order feeding f32 sums
and a stable sort, :2194,:2222) this restatement takes the same deterministic choice as the
product: query terms in bytewise order, ties by ascending chunk key.

Only tests/ may import this module."""
import math

import numpy as np

F = np.float32


def tokenize(text: str):
    """fn tokenize, :2242-2247: split on !char::is_alphanumeric, keep len() >= 3 BYTES, str::to_lowercase.
    Independent of the product's generated tables: the `regex` module's \\p{Alphabetic} / \\p{N*} classes are
    Rust's `Alphabetic || Numeric`, and CPython's str.lower() is the full lowercase mapping with Final_Sigma."""
    import regex
    toks = regex.split(r"[^\p{Alphabetic}\p{Nd}\p{Nl}\p{No}]", text)
    return [t.lower() for t in toks if len(t.encode("utf-8")) >= 3]


def logf(x):
    """f32::ln == the C library's logf (libm), which is what the product calls."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    return F(libm.logf(ctypes.c_float(float(x))))


class LexicalIndex:
    """struct LexicalIndex, :2084-2090."""

    def __init__(self):
        self.term_postings = {}
        self.doc_lengths = {}
        self.doc_terms = {}
        self.total_docs = 0
        self.total_length = 0

    def add_chunk(self, key, text):            # :2106-2138
        if key in self.doc_terms:
            self.remove_chunk(key)
        tokens = tokenize(text)
        if not tokens:
            return
        counts = {}
        for t in tokens:
            counts[t] = counts.get(t, 0) + 1
        doc_length = sum(counts.values())
        if doc_length == 0:
            return
        for term, c in counts.items():
            self.term_postings.setdefault(term, {})[key] = c
        self.doc_lengths[key] = doc_length
        self.doc_terms[key] = counts
        self.total_docs += 1
        self.total_length += doc_length

    def remove_chunk(self, key):               # :2140-2167
        counts = self.doc_terms.pop(key, None)
        if counts is not None:
            for term in counts:
                p = self.term_postings.get(term)
                if p is not None:
                    p.pop(key, None)
                    if not p:
                        del self.term_postings[term]
            length = self.doc_lengths.pop(key, None)
            if length is not None:
                self.total_length = self.total_length - length if self.total_length >= length else 0
            if self.total_docs > 0:
                self.total_docs -= 1
        else:
            self.doc_lengths.pop(key, None)
        if self.total_docs == 0:
            self.total_length = 0

    def score(self, query, limit):             # :2169-2227
        if self.total_docs == 0:
            return []
        tokens = tokenize(query)
        if not tokens:
            return []
        terms = sorted(set(tokens), key=lambda t: t.encode("utf-8"))
        avg = F(self.total_length) / F(self.total_docs)
        k1, b = F(1.5), F(0.75)
        scores = {}
        for term in terms:
            postings = self.term_postings.get(term)
            if postings is None:
                continue
            df = F(len(postings))
            idf = logf((F(self.total_docs) - df + F(0.5)) / (df + F(0.5)))
            idf = F(max(idf, F(0.0)))
            for key, tf_i in postings.items():
                dl = F(self.doc_lengths.get(key, 0))
                if dl == 0:
                    continue
                tf = F(tf_i)
                denom = F(tf + F(k1 * F(F(F(1.0) - b) + F(b * F(dl / avg)))))
                if denom == 0:
                    continue
                sc = F(F(idf * F(tf * F(k1 + F(1.0)))) / denom)
                scores[key] = F(scores.get(key, F(0.0)) + sc)
        results = sorted(scores.items(), key=lambda kv: (-float(kv[1]), kv[0]))
        if limit > 0 and len(results) > limit:
            results = results[:limit]
        return results
