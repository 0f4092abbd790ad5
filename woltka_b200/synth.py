"""Seeded synthetic workloads of BASELINE.json / SURVEY.md §8(d).

One generator feeds both worlds: int32 SoA columns for the GPU path and, for a
subsample, text (SAM + nodes.dmp + taxid map + gene coordinates) that the
unmodified reference can read — that is how tests/golden/make_golden.py builds
fixtures for modes the reference's bundled data does not exercise.

Record columns are produced with torch so the same code runs on the host and,
for the 1e8-record bench, directly in HBM.
"""
import numpy as np
import torch

RANKS = ['no rank', 'kingdom', 'phylum', 'class', 'order', 'family', 'genus',
         'species']
LEVEL_SIZES = [1, 2, 50, 150, 400, 1000, 3000, 7000]
N_GENOMES = 10000


class Taxonomy:
    """root + 7 ranked levels + genome leaves (T = 21,603 by default)."""

    def __init__(self, seed=42, level_sizes=LEVEL_SIZES, n_genomes=N_GENOMES):
        rng = np.random.default_rng(seed)
        sizes = list(level_sizes) + [n_genomes]
        level_off = np.concatenate([[0], np.cumsum(sizes)])
        T = int(level_off[-1])
        parent = np.zeros(T, dtype=np.int32)
        node_rank = np.full(T, -1, dtype=np.int32)
        for lv in range(len(sizes)):
            a, b = int(level_off[lv]), int(level_off[lv + 1])
            if lv < len(level_sizes):
                node_rank[a:b] = lv
            if lv == 0:
                continue
            # children take parents in contiguous blocks of random size, so
            # index-near genomes are taxonomically near
            pa, pb = int(level_off[lv - 1]), int(level_off[lv])
            npar, nch = pb - pa, b - a
            cuts = np.sort(rng.choice(np.arange(1, nch), npar - 1,
                                      replace=False)) if npar > 1 else []
            bounds = np.concatenate([[0], cuts, [nch]]).astype(np.int64)
            parent[a:b] = np.repeat(np.arange(pa, pb), np.diff(bounds))
        self.parent = parent
        self.node_rank = node_rank
        self.level_off = [int(x) for x in level_off]
        self.rank_names = list(RANKS[:len(level_sizes)])
        self.n_genomes = n_genomes
        self.genome_node0 = int(level_off[-2])
        self.T = T

    # identifiers the text world uses
    def node_id(self, i):
        return f'G{i - self.genome_node0:09d}' if i >= self.genome_node0 \
            else str(i + 1)

    def ids(self):
        return [self.node_id(i) for i in range(self.T)]

    def genome_id(self, g):
        return f'G{g:09d}'

    def write_nodes_dmp(self, fp):
        """NCBI nodes.dmp for the ranked part of the tree."""
        with open(fp, 'w') as f:
            for i in range(self.genome_node0):
                f.write(f'{i + 1}\t|\t{self.parent[i] + 1}\t|\t'
                        f'{self.rank_names[self.node_rank[i]]}\t|\n')

    def write_taxid_map(self, fp):
        with open(fp, 'w') as f:
            for g in range(self.n_genomes):
                f.write(f'{self.genome_id(g)}\t'
                        f'{self.parent[self.genome_node0 + g] + 1}\n')


def gen_hits(n_rec, n_genomes=N_GENOMES, n_samples=1, seed=1002,
             device='cpu', p=0.48, kmax=16, zipf=1.1, dup=0.02, window=20):
    """Alignment records of §8(d): returns (qidx, sidx, q_sample, n_qry).

    k hits per query ~ min(Geometric(p), kmax); first subject Zipf-weighted
    (per-sample rotation of the abundance ranking), further subjects within
    `window` genome indices of it, `dup` of the non-first records repeat the
    previous record exactly.  Records of a query are contiguous, queries of a
    sample are contiguous.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q_est = int(n_rec / 1.8) + 1024
    u = torch.rand(q_est, generator=g, device=dev, dtype=torch.float64)
    k = torch.floor(torch.log1p(-u) / np.log(1.0 - p)).to(torch.int64) + 1
    k.clamp_(1, kmax)
    csum = torch.cumsum(k, 0)
    n_qry = int(torch.searchsorted(csum, torch.tensor(n_rec, device=dev),
                                   right=False).item()) + 1
    k = k[:n_qry].clone()
    csum = csum[:n_qry]
    k[-1] -= csum[-1] - n_rec
    start = torch.cumsum(k, 0) - k
    qidx = torch.repeat_interleave(
        torch.arange(n_qry, device=dev, dtype=torch.int32), k)
    assert qidx.numel() == n_rec
    # Zipf abundance over genomes, rotated per sample
    w = 1.0 / torch.arange(1, n_genomes + 1, device=dev,
                           dtype=torch.float64) ** zipf
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(n_genomes, generator=g, device=dev)
    r = torch.searchsorted(cdf, torch.rand(n_qry, generator=g, device=dev,
                                           dtype=torch.float64))
    r.clamp_(max=n_genomes - 1)
    q_sample = (torch.arange(n_qry, device=dev, dtype=torch.int64) *
                n_samples // n_qry).to(torch.int32)
    first = perm[(r + q_sample.to(torch.int64) * 977) % n_genomes]
    pos = torch.arange(n_rec, device=dev, dtype=torch.int64) - \
        start[qidx.to(torch.int64)]
    off = torch.randint(0, window, (n_rec,), generator=g, device=dev)
    s = (first[qidx.to(torch.int64)] + torch.where(pos == 0, 0, off)) \
        % n_genomes
    isdup = (torch.rand(n_rec, generator=g, device=dev) < dup) & (pos > 0)
    prev = torch.roll(s, 1)
    s = torch.where(isdup, prev, s)
    return qidx, s.to(torch.int32), q_sample, n_qry


def gen_genes(n_contigs=1000, genes_per_contig=5000, contig_len=5_000_000,
              seed=1003):
    """Gene table of cfg3: per contig sorted starts, length U[300,1500)."""
    rng = np.random.default_rng(seed)
    G = n_contigs * genes_per_contig
    start = np.sort(rng.integers(1, contig_len - 1500, size=(
        n_contigs, genes_per_contig), dtype=np.int64), axis=1).reshape(-1)
    length = rng.integers(300, 1500, size=G, dtype=np.int64)
    lo, hi = start, start + length - 1          # 1-based inclusive
    contig_off = np.arange(n_contigs + 1, dtype=np.int64) * genes_per_contig
    # woltka_b200.h convention: gbeg = lo - 1, gend = hi
    return contig_off, (lo - 1).astype(np.int32), hi.astype(np.int32)


def gen_reads(n_rec, n_contigs=1000, contig_len=5_000_000, seed=1003,
              device='cpu'):
    """Reads of cfg3: (qidx, contig, beg, end, len, n_qry); 70 % single-hit
    queries, 30 % with 2-4 hits; CIGAR 150M (90 %) or 70M2D78M2S (10 %)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 1)
    q_est = int(n_rec / 1.4) + 1024
    u = torch.rand(q_est, generator=g, device=dev)
    k = torch.where(u < 0.7, 1, torch.randint(2, 5, (q_est,), generator=g,
                                              device=dev))
    csum = torch.cumsum(k, 0)
    n_qry = int(torch.searchsorted(csum, torch.tensor(n_rec, device=dev),
                                   right=False).item()) + 1
    k = k[:n_qry].clone()
    k[-1] -= csum[n_qry - 1] - n_rec
    qidx = torch.repeat_interleave(
        torch.arange(n_qry, device=dev, dtype=torch.int32), k)
    contig = torch.randint(0, n_contigs, (n_rec,), generator=g, device=dev,
                           dtype=torch.int32)
    pos = torch.randint(1, contig_len - 200, (n_rec,), generator=g,
                        device=dev, dtype=torch.int32)
    gapped = torch.rand(n_rec, generator=g, device=dev) < 0.1
    length = torch.where(gapped, 148, 150).to(torch.int32)
    beg = pos - 1
    end = beg + 150
    return qidx, contig, beg, end, length, n_qry


def write_sam(fp, qnames, subjects, pos=None, cigar=None, flags=None):
    """Minimal SAM body the reference parsers accept (align.py:258-406)."""
    with open(fp, 'w') as f:
        f.write('@HD\tVN:1.0\tSO:unsorted\n')
        for i, (qn, sn) in enumerate(zip(qnames, subjects)):
            p_ = 1 if pos is None else pos[i]
            c_ = '150M' if cigar is None else cigar[i]
            fl = 0 if flags is None else flags[i]
            f.write(f'{qn}\t{fl}\t{sn}\t{p_}\t42\t{c_}\t*\t0\t0\t*\t*\n')


def write_coords(fp, contig_off, gbeg, gend, contig_name='C%d',
                 gene_name='g%d', flip_every=2):
    """The gene table of gen_genes as a coordinates file the reference's
    load_gene_coords reads (ordinal.py:338-430): `>contig` then
    `gene<TAB>beg<TAB>end`, 1-based inclusive, every `flip_every`-th gene
    written end first (a gene on the reverse strand)."""
    with open(fp, 'w') as f:
        for c in range(len(contig_off) - 1):
            f.write('>' + contig_name % c + '\n')
            for g in range(int(contig_off[c]), int(contig_off[c + 1])):
                lo, hi = int(gbeg[g]) + 1, int(gend[g])
                if flip_every and g % flip_every == 1:
                    lo, hi = hi, lo
                f.write(f'{gene_name % (g - int(contig_off[c]))}\t{lo}\t{hi}\n')


def reads_as_sam(fp, qidx, contig, beg, length, contig_name='C%d',
                 query_name='R%d'):
    """The reads of gen_reads as SAM lines: POS = beg + 1, CIGAR 150M, or
    70M2D78M2S for the gapped ones (aligned length 148, span 150)."""
    write_sam(fp, [query_name % q for q in np.asarray(qidx).tolist()],
              [contig_name % c for c in np.asarray(contig).tolist()],
              pos=(np.asarray(beg) + 1).tolist(),
              cigar=['150M' if ln == 150 else '70M2D78M2S'
                     for ln in np.asarray(length).tolist()])


# ---- a classify problem in the integer world of include/woltka_b200.h -------
# (shared by bench.py, the GPU parity tests and smoke())
from ._lib import (KIND_NONE, KIND_FREE, KIND_RANK, F_UNIQ, F_ABOVE,  # noqa: E402
                   F_MAJOR, F_UNASSIGNED)
from .hierarchy import FlatTree  # noqa: E402


class Case:
    """A classify problem in the integer world of include/woltka_b200.h."""

    def __init__(self, tax, n_extra=0, internal_subjects=0, seed=0):
        # subjects: every genome, then `internal_subjects` internal nodes,
        # then `n_extra` subjects that are not in the tree
        rng = np.random.default_rng(seed)
        self.tax = tax
        self.ft = FlatTree.from_arrays(tax.parent, tax.node_rank,
                                       tax.rank_names, tax.level_off)
        T = tax.T
        g = np.arange(tax.n_genomes, dtype=np.int32) + tax.genome_node0
        inner = rng.integers(0, tax.genome_node0, internal_subjects,
                             dtype=np.int32)
        self.sub_node = np.concatenate(
            [g, inner, np.full(n_extra, -1, dtype=np.int32)]).astype(np.int32)
        self.sub_feat = self.sub_node.copy()
        self.sub_feat[self.sub_node < 0] = T + np.arange(n_extra)
        self.V = len(self.sub_node)
        self.NF = T + n_extra

    def tables(self, entries, subok=False):
        """entries: list of 'none' | 'free' | rank name -> (kinds, tab, trk)"""
        kinds, rows, trk = [], [], []
        for e in entries:
            if e == 'none':
                kinds.append(KIND_NONE)
                rows.append(self.sub_feat)
                trk.append(0)
            elif e == 'free':
                kinds.append(KIND_FREE)
                par = np.where(self.sub_node >= 0,
                               self.ft.parent[np.maximum(self.sub_node, 0)],
                               -1)
                rows.append(self.sub_feat if subok else par)
                trk.append(0)
            else:
                kinds.append(KIND_RANK)
                anc = self.ft.anc_at_rank(e)
                rows.append(np.where(self.sub_node >= 0,
                                     anc[np.maximum(self.sub_node, 0)], -1))
                trk.append(self.ft.rank_id(e))
        return (np.array(kinds, dtype=np.int32),
                np.stack(rows).astype(np.int32),
                np.array(trk, dtype=np.int32))



MODES = {
    'default': 0,
    'uniq': F_UNIQ,
    'above': F_ABOVE,
    'major': F_MAJOR,
    'uniq+unassigned': F_UNIQ | F_UNASSIGNED,
    'major+unassigned': F_MAJOR | F_UNASSIGNED,
    'above+unassigned': F_ABOVE | F_UNASSIGNED,
}


