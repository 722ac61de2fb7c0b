"""CPU: host-side logic of the product package (no GPU needed)."""
import os

import numpy as np

import hica_oracle as O
from hic_genome_assembler_b200 import part1, part2

G = os.path.join(os.path.dirname(__file__), "golden")


def test_enumeration_tables_match_reference_order():
    # orders documented in SURVEY.md 3.2 (measured by calling the reference)
    t3 = ["".join("ABC"[i] for i in r) for r in part2.order_table(3)]
    assert t3 == ["ABC", "ACB", "BAC"]
    t4 = ["".join("ABCD"[i] for i in r) for r in part2.order_table(4)]
    assert t4 == "ABCD ABDC ACBD ACDB ADCB ADBC BACD BADC BCAD BDAC CBAD CABD".split()
    o3 = ["".join("-" if v else "+" for v in r) for r in part2.orient_table(3)]
    assert o3 == "+++ --- +-- -+- --+ ++- +-+ -++".split()
    o4 = ["".join("-" if v else "+" for v in r) for r in part2.orient_table(4)]
    assert o4 == "++++ ---- +--- -+-- --+- ---+ ++-- +-+- +--+ -++- -+-+ --++ +++- ++-+ +-++ -+++".split()
    for n, cnt in [(5, 1920), (6, 23040), (7, 322560)]:
        assert part2.order_table(n).shape[0] * part2.orient_table(n).shape[0] == cnt == part2.calc_possible_perms(n)
    for n in range(1, 7):
        assert [tuple(int(v) for v in r) for r in part2.order_table(n)] == O.order_candidates(n)
        assert [tuple("-" if v else "+" for v in r) for r in part2.orient_table(n)] == O.orient_candidates(n)


def _read_groups(path):
    groups, cur = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("#"):
            cur = []
            groups.append(cur)
        else:
            cur.append(line.split("\t"))
    return groups


def test_chromosome_groups_match_reference_files():
    """binGroups.txt (cuts applied to the dendrogram order) -> chromosomeGroups.txt, both written by the reference."""
    d = os.path.join(G, "pipeline_320")
    bin_groups = _read_groups(os.path.join(d, "out", "binGroups.txt"))
    want = _read_groups(os.path.join(d, "out", "chromosomeGroups.txt"))
    bed = [l.rstrip("\n").split("\t") for l in open(os.path.join(d, "in", "synth_abs.bed"))]
    bin_ids = [int(c[3]) for c in bed]
    scaf = [c[0] for c in bed]
    sizes = {l.split("\t")[0]: int(l.split("\t")[1]) for l in open(os.path.join(d, "in", "synth.sizes"))}
    row_of = {b: i for i, b in enumerate(bin_ids)}
    leaves, cuts = [], []
    for g in bin_groups:
        leaves += [row_of[int(c[0])] for c in g]
        cuts.append(len(leaves))
    cuts = cuts[:-1]
    got = part1.chromosome_groups(cuts, leaves, bin_ids, scaf, sizes)
    assert [[[str(b), s] for b, s in g] for g in got] == want


def test_compact_groups_equal_tuple_groups():
    rng = np.random.RandomState(3)
    n, S = 900, 37
    code = np.sort(rng.randint(0, S, size=n))           # bins of a scaffold are contiguous in file order
    names = ["scaf_%d" % i for i in range(S)]
    sizes = [int(1000 * (code == s).sum()) for s in range(S)]
    bin_ids = rng.permutation(np.arange(1, n + 1))      # arbitrary ids
    leaves = rng.permutation(n)
    # make the dendrogram order mostly scaffold-contiguous so that groups are non-trivial
    leaves = np.argsort(code + 0.3 * rng.rand(n), kind="stable")
    cuts = [120, 121, 400, 650, 890]
    want = part1.chromosome_groups(cuts, leaves, bin_ids, [names[c] for c in code], dict(zip(names, sizes)))
    got = part1.chromosome_groups_compact(cuts, leaves, bin_ids, code, names, sizes)
    flat = [[(int(b), nm) for nm, ids in g for b in ids] for g in got]
    assert flat == want


def test_compact_groups_assignment_threshold_follows_python_round():
    """A scaffold belongs to a group when round(100 c / total, 2) >= 51 (:1060-1066).  Shares that land within
    rounding distance of the threshold (50.995 % ... 51.005 %) must be decided exactly as the reference's round()
    does, shares far from it take the vectorised path; scaffolds split 50/50 belong to no group."""
    cases = []   # (bins of the scaffold in group 0, total bins of the scaffold)
    for tot in (200, 2000, 1999, 1567, 10000, 7919):
        for c in {int(tot * .51) - 1, int(tot * .51), int(tot * .51) + 1, (tot * 50995) // 100000,
                  (tot * 50995) // 100000 + 1, (tot * 51005) // 100000, tot // 2, tot}:
            if 0 < c <= tot:
                cases.append((c, tot))
    names, code, sizes = [], [], []
    in_first = []
    for s, (c, tot) in enumerate(cases):
        names.append("s%d" % s)
        sizes.append(1000 + s)
        code += [s] * tot
        in_first += [True] * c + [False] * (tot - c)
    code = np.array(code)
    in_first = np.array(in_first)
    n = len(code)
    leaves = np.concatenate([np.nonzero(in_first)[0], np.nonzero(~in_first)[0]])
    cuts = [int(in_first.sum())]
    bin_ids = np.arange(1, n + 1)
    want = part1.chromosome_groups(cuts, leaves, bin_ids, [names[c] for c in code], dict(zip(names, sizes)))
    got = part1.chromosome_groups_compact(cuts, leaves, bin_ids, code, names, sizes)
    flat = [[(int(b), nm) for nm, ids in g for b in ids] for g in got]
    assert flat == want
    # the reference's own expression, entry by entry (the first group is the one that holds a scaffold lying in it
    # with all its bins)
    full = {names[s] for s, (c, tot) in enumerate(cases) if c == tot}
    g0 = [g for g in got if full & {nm for nm, _ in g}]
    assert len(g0) == 1
    first_group = {nm for nm, _ in g0[0]}
    for s, (c, tot) in enumerate(cases):
        assert (names[s] in first_group) == (round((float(c) / float(tot)) * 100., 2) >= 51.), (c, tot)
