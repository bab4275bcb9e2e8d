"""CPU model of the exact top-N ranking scheme of the streaming predict path ("Exactness of the ranking").

The GPU path never sorts all N running sums per read (reference src/sketchy.rs:348, 391). It (1) takes, per read, the
`top`-th best key among a small set of tracked rows as a lower bound, (2) lets the rank warps of `fused_kernel` walk
every row's per-read counters and emit INTERVALS "row i holds sum v and meets the bound for reads [b0, b1)" using a
conservative test (bounds staged at every 4th read, saturating 16-bit growth, ties decided against the largest bound
index of the lane's segment), (3) expands the intervals into per-read buckets and selects the exact top-N of each.

This file restates steps 1-3 in plain Python, lane segment by lane segment like the kernel, and checks with hypothesis
that for ANY running sums, per-read counts and tracked set: every interval carries the row's true cumulative sum for
every read it covers, every row that meets a read's exact bound is in that read's bucket, and the selection equals
the stable descending sort of the reference. It is a model of the algorithm, not of the CUDA code (the `-m gpu` tests
check the code); what it pins down is that the conservative shortcuts can only add candidates, never lose one."""
from hypothesis import given, settings, strategies as st

LANES = 4          # lanes of the model "warp"
SAT = 7            # saturation value of the staged bound growth (0xFFFF in the kernel)


def better(sa, ia, sb, ib):
    """(sum desc, index asc): the order of the reference's stable descending sort."""
    return sa > sb or (sa == sb and ia < ib)


def exact_bounds(cum, tracked, top):
    """per read: the top-th best key among the tracked rows (their exact sums at that read)"""
    out = []
    keep = min(top, len(tracked))
    for b in range(len(cum[0])):
        keys = sorted(((-cum[t][b], t) for t in tracked))
        s, i = keys[keep - 1]
        out.append((-s, i))
    return out


def rank_warp_row(carry, counts_row, gi, lb, per, emit):
    """One row through the rank warp (kernels_predict.cu, rank warps): lanes own `per` consecutive reads."""
    n_reads = len(counts_row)
    lb_min = lb[0][0]
    lbrel = [min(lb[4 * i][0] - lb_min, SAT) for i in range((n_reads + 3) // 4)]

    def is_cand(sv, b, li_cap):
        rel = lbrel[b >> 2]
        ls = lb_min + rel if rel != SAT else lb[b & ~3][0]   # saturated: the kernel reads the bound itself
        return sv > ls or (sv == ls and gi <= li_cap)

    row_total = sum(counts_row)
    if row_total:
        if carry + row_total < lb_min:
            return
        incl = 0
        for lane in range(LANES):
            seg0 = lane * per
            seg = counts_row[seg0:seg0 + per]
            tot = sum(seg)
            incl += tot
            if seg0 >= n_reads:
                continue
            lb_seg = lb_min + lbrel[seg0 >> 2]   # saturates like the staged growth: lower, so only more permissive
            if carry + incl < lb_seg:
                continue
            seg_end = min(seg0 + per, n_reads)
            li_seg = max(lb[b][1] for b in range(seg0, seg_end))
            run = incl - tot
            is_open, first, ob, osum = False, True, 0, 0
            for w0 in range(seg0, seg_end, 4):           # one 32-bit word = four u8 counters
                word = counts_row[w0:w0 + 4]
                if not any(word) and not is_open and not first:
                    continue
                for b in range(w0, min(w0 + 4, seg_end)):
                    c = counts_row[b]
                    check = is_open or first
                    first = False
                    if c:
                        if is_open:
                            emit(osum, gi, ob, b)
                            is_open = False
                        run += c
                        check = True
                    if check:
                        sv = carry + run
                        cand = sv >= lb_seg and is_cand(sv, b, li_seg)
                        if cand and not is_open:
                            is_open, ob, osum = True, b, sv
                        if not cand and is_open:
                            emit(osum, gi, ob, b)
                            is_open = False
            if is_open:
                emit(osum, gi, ob, seg_end)
    elif carry >= lb_min:
        li_all = max(i for _, i in lb)
        e = 0
        while e < n_reads and is_cand(carry, e, li_all):
            e += 1
        if e:
            emit(carry, gi, 0, e)


@st.composite
def worlds(draw):
    n_rows = draw(st.integers(1, 9))
    per = draw(st.sampled_from([4, 8]))
    n_reads = draw(st.integers(1, LANES * per))
    top = draw(st.integers(1, 4))
    sums0 = [draw(st.integers(0, 6)) for _ in range(n_rows)]
    counts = [[draw(st.sampled_from([0, 0, 0, 1, 2, 5])) for _ in range(n_reads)] for _ in range(n_rows)]
    n_tr = draw(st.integers(min(top, n_rows), n_rows))    # the library always tracks >= min(top, n_rows) distinct rows
    tracked = draw(st.permutations(list(range(n_rows))))[:n_tr]
    return n_rows, per, n_reads, top, sums0, counts, tracked


@settings(max_examples=600, deadline=None)
@given(worlds())
def test_intervals_cover_every_true_candidate_and_selection_is_exact(w):
    n_rows, per, n_reads, top, sums0, counts, tracked = w
    cum = [[sums0[i] + sum(counts[i][:b + 1]) for b in range(n_reads)] for i in range(n_rows)]
    lb = exact_bounds(cum, tracked, top)
    assert all(lb[b][0] <= lb[b + 1][0] for b in range(n_reads - 1))   # sums never decrease, so neither do bounds
    buckets = [[] for _ in range(n_reads)]

    def emit(sv, gi, b0, b1):
        assert 0 <= b0 < b1 <= n_reads
        for b in range(b0, b1):
            assert cum[gi][b] == sv, "an interval must carry the row's sum at every read it covers"
            buckets[b].append((sv, gi))

    for i in range(n_rows):
        rank_warp_row(sums0[i], counts[i], i, lb, per, emit)
    for b in range(n_reads):
        rows_in = [gi for _, gi in buckets[b]]
        assert len(rows_in) == len(set(rows_in)), "a row enters a read's bucket at most once"
        bs, bi = lb[b]
        for i in range(n_rows):
            if not better(bs, bi, cum[i][b], i):   # key at least as good as the bound: a true candidate
                assert i in rows_in, (b, i)
        want = sorted(range(n_rows), key=lambda i: (-cum[i][b], i))[:top]
        got = sorted(buckets[b], key=lambda c: (-c[0], c[1]))[:top]
        # the bucket holds >= min(top, n_rows) rows (the tracked rows that define the bound are candidates themselves)
        assert [gi for _, gi in got] == want[:len(got)] and len(got) == min(top, n_rows)
        assert [sv for sv, _ in got] == [cum[i][b] for i in want]
