"""CPU: the rule that lets the grouped fp32 leaf kernels (csrc/gemm_f32.cu: k_g32_merge -> k_gemm_f32_g32 / g64) compute a 2x2
group of C tiles per MMA without changing the executed-product set.  Per contraction index k the executed combinations
W of a group are a subset of {0,1} x {0,1} (member m = 2c + r).  A super-product is (A rows used, B columns used) with the
others replaced by a zero tile; it contributes A_r * B_c to member (r, c) for every used r and c.  The rule: one
super-product if W is a rectangle, else one per row.  Enumerated here for all 16 subsets: every wanted combination is
computed exactly once and no unwanted combination gets a non-zero contribution."""
import itertools


def super_products(w):
    """Python restatement of the decision in k_g32_merge (w[m] for m = 2c + r): list of (rows_used, cols_used)."""
    nw = sum(w)
    if nw == 0:
        return []
    rows = {m & 1 for m in range(4) if w[m]}
    cols = {m >> 1 for m in range(4) if w[m]}
    if nw == len(rows) * len(cols):
        return [(rows, cols)]
    return [({0}, {c for c in (0, 1) if w[2 * c + 0]}), ({1}, {c for c in (0, 1) if w[2 * c + 1]})]


def test_rectangle_rule_is_exact_for_every_subset():
    for w in itertools.product((False, True), repeat=4):
        got = {}
        for rows, cols in super_products(w):
            assert rows and cols                      # never an empty super-product
            for r in rows:
                for c in cols:
                    got[2 * c + r] = got.get(2 * c + r, 0) + 1
        want = {m: 1 for m in range(4) if w[m]}
        assert got == want, (w, got)


def test_rule_uses_the_minimum_number_of_mmas():
    """1 super-product whenever one suffices (W a rectangle), never more than 2."""
    for w in itertools.product((False, True), repeat=4):
        sp = super_products(w)
        rows = {m & 1 for m in range(4) if w[m]}
        cols = {m >> 1 for m in range(4) if w[m]}
        is_rect = sum(w) == len(rows) * len(cols)
        assert len(sp) == (0 if not any(w) else 1 if is_rect else 2)


def test_merged_k_list_keeps_every_members_own_order():
    """The group's super-products are emitted in ascending k (a 4-way merge of the members' ascending k-lists), so each
    member accumulates its own k-list in its own order -- interleaved with exact zeros from the k's it does not have."""
    import random
    rng = random.Random(5)
    for _ in range(200):
        lists = [sorted(rng.sample(range(40), rng.randint(0, 12))) for _ in range(4)]
        ptr = [0, 0, 0, 0]
        seen = [[] for _ in range(4)]
        while True:
            heads = [lists[m][ptr[m]] for m in range(4) if ptr[m] < len(lists[m])]
            if not heads:
                break
            k = min(heads)
            w = [ptr[m] < len(lists[m]) and lists[m][ptr[m]] == k for m in range(4)]
            for rows, cols in super_products(w):
                for r in rows:
                    for c in cols:
                        seen[2 * c + r].append(k)
            for m in range(4):
                if w[m]:
                    ptr[m] += 1
        assert seen == lists
