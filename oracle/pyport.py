"""Pure-Python restatement of the reference's classify path, in the reference's
own world of str / set / dict.

TEST INFRASTRUCTURE ONLY (see oracle/woltka_oracle.c for the rules).  Two uses:
  * a second, string-level checker: it keeps the reference's chunking and
    summation scheme (per-chunk dict added into the running dict); its sums
    differ from the reference's only through the hash order of Python sets
    (tests/test_pyport.py checks it on the golden fixtures to 1e-9);
  * the "pure-Python CPU path" figure bench.py reports next to the C port:
    same interpreter-bound data structures as the reference, so the same
    order of magnitude of speed.

Every function cites the reference lines it follows
(/root/reference/woltka/).
"""
from collections import defaultdict


# ---- tree.py --------------------------------------------------------------
def ancestor_at_rank(taxon, rank, tree, rankdic):
    """tree.find_rank (tree.py:467-510)."""
    if taxon not in tree:
        return None
    node = taxon
    while True:
        if rankdic.get(node) == rank:
            return node
        up = tree[node]
        if up == node:
            return None
        node = up


def lineage(taxon, tree):
    """tree.get_lineage (tree.py:391-432): root first."""
    if taxon not in tree:
        return None
    path = [taxon]
    while tree[path[-1]] != path[-1]:
        path.append(tree[path[-1]])
    path.reverse()
    return path


def lowest_common_ancestor(taxa, tree):
    """tree.find_lca (tree.py:513-566)."""
    it = iter(taxa)
    shared = lineage(next(it), tree)
    if shared is None:
        return None
    for taxon in it:
        if taxon not in tree:
            return None
        node = taxon
        while True:
            if node in shared:
                del shared[shared.index(node) + 1:]
                break
            if tree[node] == node:
                break
            node = tree[node]
    return shared[-1]


# ---- classify.py ------------------------------------------------------------
def top_by_majority(taxa, th):
    """classify.majority (classify.py:300-317) + util.count_list."""
    counts = {}
    for t in taxa:
        counts[t] = counts.get(t, 0) + 1
    best = max(counts.values())
    for t, n in counts.items():      # first seen among the top counts
        if n == best:
            return t if n >= len(taxa) * th else None


def assign(subs, kind, rank, tree, rankdic, root, uniq, major, above, subok):
    """classify.assign_none / assign_free / assign_rank (classify.py:32-127)."""
    if kind == 'none':
        if len(subs) == 1:
            return subs[0]
        return None if uniq else list(subs)
    if kind == 'free':
        if len(subs) == 1:
            sub = subs[0]
            return sub if subok else tree.get(sub)
        lca = lowest_common_ancestor(subs, tree)
        return None if lca == root else lca
    taxa = [ancestor_at_rank(s, rank, tree, rankdic) for s in subs]
    distinct = set(taxa)
    if len(distinct) == 1:
        return taxa[0]
    if major:
        return top_by_majority(taxa, major)
    if above:
        if None in distinct:
            return None
        lca = lowest_common_ancestor(distinct, tree)
        return None if lca == root else lca
    if uniq:
        return None
    return taxa


def tally(queries, results, strata=None):
    """classify.counter / counter_strat (classify.py:144-171, 216-249)."""
    res = defaultdict(int)
    for query, taxa in zip(queries, results):
        if not taxa:
            continue
        if strata is not None:
            if query not in strata:
                continue
            key = (lambda t, s=strata[query]: (s, t))
        else:
            key = (lambda t: t)
        if isinstance(taxa, list):
            kept = [t for t in taxa if t]
            share = 1 / len(kept)
            for t in kept:
                res[key(t)] += share
        else:
            res[key(taxa)] += 1
    return res


def tally_sized(subjects, results, sizes, queries=None, strata=None):
    """classify.counter_size / counter_size_strat (classify.py:174-213,
    252-297): a uniquely assigned query adds the mean weight of its subjects,
    a list adds weight / k' per listed subject; a subject without a weight is
    a KeyError.  With strata only the queries that have one count, keyed by
    (stratum, taxon)."""
    res = defaultdict(int)
    for i, (subs, taxa) in enumerate(zip(subjects, results)):
        if not taxa:
            continue
        if strata is not None:
            if queries[i] not in strata:
                continue
            key = (lambda t, s=strata[queries[i]]: (s, t))
        else:
            key = (lambda t: t)
        if isinstance(taxa, list):
            share = 1 / len([t for t in taxa if t])
            for t, sub in zip(taxa, subs):
                if t:
                    res[key(t)] += sizes[sub] * share
        else:
            res[key(taxa)] += sum(sizes[x] for x in subs) / len(subs)
    return res


# ---- workflow.py ---------------------------------------------------------------
def split_by_sample(qryque, subque, samples=None):
    """workflow.demultiplex (workflow.py:844-909)."""
    out = {}
    for query, subs in zip(qryque, subque):
        left, _, right = query.partition('_')
        sample, read = (left, right) if right else ('', left)
        if samples and sample not in samples:
            continue
        qs, ss = out.setdefault(sample, ([], []))
        qs.append(read)
        ss.append(subs)
    return out


def readmap_lines(queries, results, namedic=None):
    """file.write_readmap (file.py:469-500) as a list of lines."""
    label = (lambda t: namedic[t] if namedic and t in namedic else t)
    out = []
    for query, taxa in zip(queries, results):
        if not taxa:
            continue
        if isinstance(taxa, list):
            counts = {}
            for t in taxa:
                if t:
                    counts[t] = counts.get(t, 0) + 1
            ranked = sorted(counts.items(), key=lambda x: (-x[1], x[0]))
            out.append('\t'.join([query] + [f'{label(t)}:{c}'
                                             for t, c in ranked]))
        else:
            out.append(f'{query}\t{label(taxa)}')
    return out


def classify_chunks(chunks, ranks, tree=None, rankdic=None, root=None,
                    uniq=False, major=None, above=False, subok=False,
                    unasgd=False, demux=False, samples=None, sample=None,
                    trimsub=None, strata_of=None, maps=None, namedic=None,
                    sizes=None):
    """workflow.classify body (workflow.py:304-335, 1017-1058) over an
    iterable of (qryque, subque) chunks; `major` is the fraction."""
    data = {r: {} for r in ranks}
    for qryque, subque in chunks:
        groups = split_by_sample(qryque, subque, samples) if demux else \
            {sample: (qryque, subque)}
        for sname, (qs, ss) in groups.items():
            if trimsub:
                ss = [{x.rsplit(trimsub, 1)[0] for x in subs} for subs in ss]
            ss = [tuple(subs) for subs in ss]
            strata = strata_of(sname) if strata_of else None
            for rank in ranks:
                kind = 'none' if rank is None or rank == 'none' or \
                    tree is None else ('free' if rank == 'free' else 'rank')
                res = [assign(s, kind, rank, tree, rankdic, root, uniq, major,
                              above, subok) for s in ss]
                if unasgd:
                    res = [x or 'Unassigned' for x in res]
                if maps is not None:
                    maps.setdefault(rank, {}).setdefault(sname, []).extend(
                        readmap_lines(qs, res, namedic))
                counts = tally(qs, res, strata) if sizes is None else \
                    tally_sized(ss, res, sizes, qs, strata)
                total = data[rank].setdefault(sname, {})
                for k, v in counts.items():           # util.sum_dict
                    total[k] = total.get(k, 0) + v
    return data


def merge_ranges(ranges):
    """range.merge_ranges (range.py:79-109): sort the (start, end) pairs of an
    interleaved list and fuse those that overlap or touch (`cend >= start`)."""
    res = []
    cstart = cend = None
    for start, end in sorted(zip(ranges[0::2], ranges[1::2])):
        if cend is None:
            cstart, cend = start, end
        elif cend >= start:
            cend = max(cend, end)
        else:
            res.extend((cstart, cend))
            cstart, cend = start, end
    if cend is not None:
        res.extend((cstart, cend))
    return res
