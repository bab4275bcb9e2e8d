"""Minimal independent Cap'n Proto codec (Python) used ONLY by the tests to cross-check the C++ .msh reader/writer:
a reader that follows struct / list / far / double-far pointers and a builder that can scatter objects over several
segments so the C++ reader's far-pointer handling is exercised. Mash `MinHash` layout as in sketchy_b200/host/msh.hpp."""
import struct


class Msg:
    def __init__(self, data: bytes):
        nseg = struct.unpack_from("<I", data, 0)[0] + 1
        sizes = struct.unpack_from("<%dI" % nseg, data, 4)
        at = 4 + 4 * nseg
        if at % 8:
            at += 4
        self.segs = []
        for s in sizes:
            self.segs.append(struct.unpack_from("<%dQ" % s, data, at))
            at += 8 * s

    def follow(self, seg, w):
        p = self.segs[seg][w]
        if p == 0:
            return None
        kind = p & 3
        if kind == 2:
            dbl, off, tseg = (p >> 2) & 1, (p >> 3) & 0x1FFFFFFF, p >> 32
            if not dbl:
                pad = self.segs[tseg][off]
                o = ((pad & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 >> 2
                return tseg, off + 1 + o, pad
            far2, tag = self.segs[tseg][off], self.segs[tseg][off + 1]
            return far2 >> 32, (far2 >> 3) & 0x1FFFFFFF, tag
        o = ((p & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000 >> 2
        return seg, w + 1 + o, p

    def struct_at(self, loc):
        seg, w, d = loc
        return seg, w, (d >> 32) & 0xFFFF, (d >> 48) & 0xFFFF

    def list_at(self, loc):
        seg, w, d = loc
        es, n = (d >> 32) & 7, d >> 35
        if es == 7:
            tag = self.segs[seg][w]
            return seg, w + 1, es, (tag & 0xFFFFFFFF) >> 2, (tag >> 32) & 0xFFFF, (tag >> 48) & 0xFFFF
        return seg, w, es, n, 0, 0

    def prims(self, loc, nbytes):
        if loc is None:
            return []
        seg, w, es, n, _, _ = self.list_at(loc)
        raw = struct.pack("<%dQ" % (len(self.segs[seg]) - w), *self.segs[seg][w:])
        fmt = {1: "B", 4: "I", 8: "Q"}[nbytes]
        return list(struct.unpack_from("<%d%s" % (n, fmt), raw, 0))

    def text(self, loc):
        b = bytes(self.prims(loc, 1))
        return b[:-1].decode() if b else ""


def decode_msh(data: bytes):
    m = Msg(data)
    seg, w, dw, pw = m.struct_at(m.follow(0, 0))
    d = m.segs[seg][w:w + dw]
    out = {"k": d[0] & 0xFFFFFFFF, "s": d[1] & 0xFFFFFFFF, "seed": ((d[2] >> 32) & 0xFFFFFFFF) ^ 42, "sketches": []}
    rl = m.follow(seg, w + dw + 3) or m.follow(seg, w + dw + 0)
    lseg, lw, ldw, lpw = m.struct_at(rl)
    eseg, ew, es, n, edw, epw = m.list_at(m.follow(lseg, lw + ldw))
    for i in range(n):
        b = ew + i * (edw + epw)
        dd = m.segs[eseg][b:b + edw]
        p = b + edw
        out["sketches"].append({
            "name": m.text(m.follow(eseg, p + 2)), "comment": m.text(m.follow(eseg, p + 3)),
            "seq_length": dd[1], "num_valid_kmers": dd[2],
            "hashes": m.prims(m.follow(eseg, p + 5), 8), "counts": m.prims(m.follow(eseg, p + 6), 4)})
    return out


def decode_msh_header(data: bytes) -> dict:
    """The file-level fields next to k / s / seed: windowSize, concatenated, alphabet (what Mash tooling reads and finch's
    writer fills)."""
    m = Msg(data)
    seg, w, dw, pw = m.struct_at(m.follow(0, 0))
    d = m.segs[seg][w:w + dw]
    return {"window": (d[0] >> 32) & 0xFFFFFFFF, "concatenated": bool((d[1] >> 32) & 1), "noncanonical": bool((d[1] >> 33) & 1),
            "alphabet": m.text(m.follow(seg, w + dw + 2))}


def encode_msh_multiseg(f: dict) -> bytes:
    """Segment 0 holds only a FAR root pointer; the root struct lives in segment 1; every sketch's hash list lives in
    its own segment behind a far pointer; names use DOUBLE-far pointers (landing pads in segment 2)."""
    segs = [[0], [], []]  # seg0: root far ptr; seg1: structs; seg2: landing pads for double-far

    def alloc(seg, n):
        at = len(segs[seg])
        segs[seg].extend([0] * n)
        return at

    def struct_ptr(ptr_at, target, dw, pw):
        return (((target - ptr_at - 1) << 2) & 0xFFFFFFFF) | (dw << 32) | (pw << 48)

    def list_ptr(ptr_at, target, es, n):
        return ((((target - ptr_at - 1) << 2) & 0xFFFFFFFF) | 1) | (es << 32) | (n << 35)

    def put_bytes(seg, at, b):
        b = b + b"\0" * (-len(b) % 8)
        for i in range(len(b) // 8):
            segs[seg][at + i] = struct.unpack_from("<Q", b, 8 * i)[0]

    n = len(f["sketches"])
    pad = alloc(1, 1)                       # landing pad of the root far pointer
    root = alloc(1, 7)
    segs[1][pad] = struct_ptr(pad, root, 3, 4)
    segs[0][0] = 2 | (pad << 3) | (1 << 32)
    segs[1][root + 0] = f["k"]
    segs[1][root + 1] = f["s"]
    segs[1][root + 2] = ((f["seed"] ^ 42) & 0xFFFFFFFF) << 32
    plist = alloc(1, 1)
    segs[1][root + 3 + 3] = struct_ptr(root + 6, plist, 0, 1)
    tag = alloc(1, 1 + 10 * n)
    segs[1][tag] = (n << 2) | (3 << 32) | (7 << 48)
    segs[1][plist] = list_ptr(plist, tag, 7, 10 * n)
    for i, s in enumerate(f["sketches"]):
        e = tag + 1 + 10 * i
        segs[1][e + 0] = min(s["seq_length"], 0xFFFFFFFF)
        segs[1][e + 1] = s["seq_length"]
        segs[1][e + 2] = s["num_valid_kmers"]
        # name: double-far -> content in a fresh segment, landing pad (far + tag) in segment 2
        nb = s["name"].encode() + b"\0"
        cseg = len(segs)
        segs.append([0] * ((len(nb) + 7) // 8))
        put_bytes(cseg, 0, nb)
        lp = alloc(2, 2)
        segs[2][lp] = 2 | (0 << 3) | (cseg << 32)                    # far pointer to content start
        segs[2][lp + 1] = 1 | (2 << 32) | (len(nb) << 35)           # tag: byte list
        segs[1][e + 3 + 2] = 2 | (1 << 2) | (lp << 3) | (2 << 32)    # double-far
        # comment: plain near text
        cb = s["comment"].encode() + b"\0"
        ct = alloc(1, (len(cb) + 7) // 8)
        put_bytes(1, ct, cb)
        segs[1][e + 3 + 3] = list_ptr(e + 6, ct, 2, len(cb))
        # hashes64: own segment behind a single far pointer (landing pad = list pointer in that segment)
        hseg = len(segs)
        segs.append([0] + list(s["hashes"]))
        segs[hseg][0] = list_ptr(0, 1, 5, len(s["hashes"]))
        segs[1][e + 3 + 5] = 2 | (0 << 3) | (hseg << 32)
        # counts32: near
        cw = alloc(1, (len(s["counts"]) * 4 + 7) // 8)
        put_bytes(1, cw, struct.pack("<%dI" % len(s["counts"]), *s["counts"]))
        segs[1][e + 3 + 6] = list_ptr(e + 9, cw, 4, len(s["counts"]))
    out = struct.pack("<I", len(segs) - 1) + struct.pack("<%dI" % len(segs), *[len(x) for x in segs])
    if len(out) % 8:
        out += b"\0" * 4
    for x in segs:
        out += struct.pack("<%dQ" % len(x), *x)
    return out
