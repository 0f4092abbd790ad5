#!/usr/bin/env python
"""Generate tests/golden/ by running the UNMODIFIED reference
(/root/reference, woltka 0.1.7) in the build container.

    python tests/golden/make_golden.py

For every case it stores the inputs the product needs (alignment files copied
or reduced from the reference's bundled test data, or produced by the seeded
synthetic generator; the hierarchy as the dicts `build_hierarchy` returned,
restricted to what the case can reach) and the reference's answer: the raw
dict `woltka.workflow.classify()` returned and the same after
`round_profiles`.  The GPU box has neither /root/reference nor biom-format,
so tests only read these fixtures.

`biom-format` is not installed here; `woltka.workflow` imports it for the
table writers only, so a three-line stub module is put on sys.path.
"""
import bz2
import copy
import gzip
import json
import lzma
import os
import shutil
import sys
import tempfile
from os.path import join, dirname, abspath, basename

import numpy as np

HERE = dirname(abspath(__file__))
ROOT = dirname(dirname(HERE))
REF = '/root/reference'
DATA = join(REF, 'woltka', 'tests', 'data')
OUT = join(HERE, 'data')
sys.path.insert(0, ROOT)

stub = tempfile.mkdtemp()
os.makedirs(join(stub, 'biom'))
with open(join(stub, 'biom', '__init__.py'), 'w') as f:
    f.write('class Table: pass\n\ndef load_table(*a, **k): raise IOError\n')
with open(join(stub, 'biom', 'util.py'), 'w') as f:
    f.write('def biom_open(*a, **k): raise IOError\n')
sys.path.insert(0, stub)
sys.path.insert(0, REF)
os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True

from woltka import workflow as W          # noqa: E402
from woltka.file import readzip           # noqa: E402
from woltka_b200 import synth             # noqa: E402


def openany(fp, mode='rt'):
    if fp.endswith('.xz'):
        return lzma.open(fp, mode)
    if fp.endswith('.bz2'):
        return bz2.open(fp, mode)
    if fp.endswith('.gz'):
        return gzip.open(fp, mode)
    return open(fp, mode.replace('t', ''))


def reduce_sam(src, dst):
    """Keep the six columns the readers use (align.py:376); drop the rest."""
    with openany(src) as fi, lzma.open(dst, 'wt') as fo:
        for line in fi:
            if line[0] == '@':
                fo.write(line)
                continue
            x = line.rstrip('\n').split('\t')
            fo.write('\t'.join(x[:6] + ['*', '0', '0', '*', '*']) + '\n')


def copy_dir(src, dst, reduce=False):
    if ONLY is not None and os.path.isdir(dst):
        return
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        sp = join(src, fn)
        if os.path.isdir(sp):
            continue
        if reduce and '.sam' in fn:
            reduce_sam(sp, join(dst, fn))
        else:
            shutil.copyfile(sp, join(dst, fn))


def closure(tree, seeds):
    """Sub-tree holding every seed and all of its ancestors."""
    keep = {}
    for s in seeds:
        cur = s
        while cur in tree and cur not in keep:
            keep[cur] = tree[cur]
            if tree[cur] == cur or tree[cur] is None:
                break
            cur = tree[cur]
    return keep


def enc_key(k):
    if isinstance(k, tuple):
        return '\x1f'.join(k)
    return '\x00' if k is None else k


def enc(data):
    return {rank: {enc_key(s): {enc_key(f): v for f, v in prof.items()}
                   for s, prof in samples.items()}
            for rank, samples in data.items()}


def subjects_of(files, fmt=None, trimsub=None, mapper=None):
    """All subject ids a plain alignment set can produce."""
    from woltka.align import plain_mapper
    subs = set()
    for fp in files:
        with readzip(fp, {}) as fh:
            for _, subque in plain_mapper(fh, fmt=fmt, n=65536):
                for ss in subque:
                    subs.update(ss)
    if trimsub:
        subs = {x.rsplit(trimsub, 1)[0] for x in subs}
    return subs


CASES = []
# `make_golden.py --only name[,name...]`: add cases to an existing tests/golden/
# (the bundled inputs under data/ are kept, other fixtures are not rewritten)
ONLY = None


def run_case(name, input_rel, *, hier=None, ranks, coords_rel=None,
             overlap=80, strata_rel=None, fmt=None, demux=None, samples=None,
             trimsub=None, uniq=False, major=None, above=False, subok=False,
             unasgd=False, exclude=None, chunk=None, note='', maps=False,
             name_as_id=False, sizes=None, cov=None):
    if ONLY is not None and name not in ONLY:
        return
    input_fp = join(OUT, input_rel)
    samples_, files, demux_ = W.parse_samples(input_fp, None, samples, demux)
    tree = rankdic = namedic = root = None
    if hier:
        tree, rankdic, namedic, root = W.build_hierarchy(**hier)
    covdir = tempfile.mkdtemp() if cov else None
    mapper, chunk_ = W.build_mapper(
        join(OUT, coords_rel) if coords_rel else None, covdir, overlap, chunk,
        {})
    ranks_, _ = W.prepare_ranks(ranks, None, tree, rankdic)
    stratmap = W.parse_strata(join(OUT, strata_rel), samples_) \
        if strata_rel else None
    excl = W.parse_exclude(exclude)
    sizemap = W.parse_sizes(sizes if sizes in (None, '.') else join(OUT, sizes),
                            mapper, {})
    rank2dir = None
    if maps:
        mapdir = tempfile.mkdtemp()
        rank2dir = {}
        for r in ranks_:
            rank2dir[r] = join(mapdir, str(r))
            os.makedirs(rank2dir[r])
    data = W.classify(mapper, files, samples_, fmt, demux_, trimsub, tree,
                      rankdic, namedic if name_as_id else None, root, ranks_,
                      rank2dir, None, uniq, major, above, subok, sizemap, unasgd,
                      stratmap, excl, chunk_, 1024, {}, covdir,
                      None if cov in (None, True) else cov)
    expected_cov = None
    if cov:
        expected_cov = {}
        for fn in sorted(os.listdir(covdir)):
            with open(join(covdir, fn)) as fh:
                expected_cov[fn] = fh.read()
        shutil.rmtree(covdir)
    expected_maps = None
    if maps:
        expected_maps = {}
        for r, d in rank2dir.items():
            expected_maps[str(r)] = {}
            for fn in sorted(os.listdir(d)):
                with open(join(d, fn)) as fh:
                    expected_maps[str(r)][fn[:-4]] = fh.read().splitlines()
        shutil.rmtree(mapdir)
    raw = copy.deepcopy(data)
    W.round_profiles(data, None)
    # hierarchy restricted to what this case can reach
    tree_small = None
    if tree is not None:
        seeds = set()
        for rank in raw.values():
            for prof in rank.values():
                for k in prof:
                    seeds.update(k if isinstance(k, tuple) else (k,))
        if coords_rel:
            # every gene a read can match: take the subjects of a rank-none run
            d2 = W.classify(mapper, files, samples_, fmt, demux_, trimsub,
                            None, None, None, None, ['none'], None, None,
                            False, None, False, False, None, False, None,
                            excl, chunk_, 1024, {}, None, None)
            for prof in d2['none'].values():
                seeds.update(prof)
        else:
            flist = files if isinstance(files, list) else list(files)
            seeds.update(subjects_of(flist, fmt, trimsub))
        tree_small = closure(tree, seeds)
    prefix = mapper.keywords['prefix'] if coords_rel else None
    if sizemap:
        # only the subjects this case can reach (a run without a hierarchy
        # lists them as its features)
        d3 = W.classify(mapper, files, samples_, fmt, demux_, trimsub, None,
                        None, None, None, ['none'], None, None, False, None,
                        False, False, None, False, None, excl, chunk_, 1024,
                        {}, None, None)
        reach = set()
        for prof in d3['none'].values():
            reach.update(prof)
        sizemap = {k: v for k, v in sizemap.items() if k in reach}
    case = {
        'name': name, 'note': note, 'input': input_rel,
        'files': ({basename(k): v for k, v in files.items()}
                  if isinstance(files, dict) else
                  [basename(x) for x in files]),
        'samples': samples_, 'demux': demux_, 'fmt': fmt, 'trimsub': trimsub,
        'tree': tree_small,
        'rankdic': ({k: v for k, v in rankdic.items() if k in tree_small}
                    if tree_small is not None and rankdic else rankdic),
        'root': root, 'ranks': ranks_, 'uniq': uniq, 'major': major,
        'above': above, 'subok': subok, 'unasgd': unasgd,
        'exclude': sorted(excl) if excl else None, 'coords': coords_rel,
        'overlap': overlap, 'prefix': prefix, 'strata': strata_rel,
        'chunk': chunk, 'expected_maps': expected_maps, 'sizes': sizemap,
        'cov': cov, 'expected_cov': expected_cov,
        'namedic': ({k: v for k, v in namedic.items() if k in tree_small}
                    if name_as_id and tree_small is not None else None),
        'expected_raw': enc(raw),
        'expected_rounded': enc(data),
    }
    with open(join(HERE, f'{name}.json'), 'w') as f:
        json.dump(case, f, sort_keys=True)
    ncell = sum(len(p) for r in data.values() for p in r.values())
    print(f'{name}: {ncell} cells')
    CASES.append(name)


def bundled():
    tax = join(DATA, 'taxonomy')
    fun = join(DATA, 'function')
    copy_dir(join(DATA, 'align', 'bowtie2'), join(OUT, 'bowtie2'), True)
    copy_dir(join(DATA, 'align', 'bt2sho'), join(OUT, 'bt2sho'), True)
    copy_dir(join(DATA, 'align', 'blastn'), join(OUT, 'blastn'))
    copy_dir(join(DATA, 'align', 'burst'), join(OUT, 'burst'))
    copy_dir(join(DATA, 'align', 'burst', 'split'), join(OUT, 'split'))
    copy_dir(join(DATA, 'output', 'burst.genus.map'),
             join(OUT, 'burst.genus.map'))
    nodes = dict(nodes_fps=[join(tax, 'nodes.dmp')],
                 map_fps=[join(tax, 'taxid.map')],
                 names_fps=[join(tax, 'names.dmp')])

    # cfg1 of BASELINE.json; tests/test_cli.py:50-53, test_workflow.py:38-48
    run_case('bowtie2_ogu', 'bowtie2', ranks=None,
             note='woltka classify -i align/bowtie2 (bowtie2.ogu.tsv)')
    run_case('bowtie2_free', 'bowtie2', hier=nodes, ranks='free',
             note='bowtie2.free.tsv (tests/test_cli.py:55-61)')
    run_case('blastn_species', join('blastn', 'mux.b6o.xz'),
             hier=dict(lineage_fps=[join(tax, 'lineages.txt')]),
             ranks='species',
             note='blastn.species.tsv: demux + 22.5 % multi-hit fractions')
    run_case('burst_genus', 'burst', hier=nodes, ranks='genus',
             note='burst.genus.tsv (tests/test_cli.py:70-79)')
    run_case('blastn_family', join('blastn', 'mux.b6o.xz'), hier=nodes,
             ranks='family', note='blastn.family.percent.tsv before --frac')
    run_case('bt2sho_phylo', 'bt2sho',
             hier=dict(newick_fps=[join(DATA, 'tree.nwk')]), ranks='free',
             subok=True, note='bt2sho.phylo.tsv: newick, free, --subok')
    run_case('bt2sho_filt_ogu', 'bt2sho', ranks=None, exclude='G000215745',
             note='bt2sho.filt.ogu.tsv: --exclude')
    run_case('split_genus', 'split',
             hier=dict(nodes_fps=[join(tax, 'nodes.dmp')],
                       map_fps=[join(tax, 'nucl', 'nucl2tid.txt')],
                       names_fps=[join(tax, 'names.dmp')]),
             ranks='genus', trimsub='_', note='split.genus.tsv: --trim-sub')
    # modes the bundled goldens do not cover, on bundled multi-hit data
    for mode, kw in [('uniq', dict(uniq=True)), ('major80', dict(major=80)),
                     ('major54', dict(major=54)), ('above', dict(above=True)),
                     ('unasgd', dict(unasgd=True, uniq=True))]:
        run_case(f'blastn_multi_{mode}', join('blastn', 'mux.b6o.xz'),
                 hier=nodes, ranks='phylum,genus,species,none,free', **kw,
                 note='multi-rank in one run')
    # read maps (tests/test_cli.py:70-90 checks burst.genus.map/*.txt.gz)
    run_case('burst_genus_map', 'burst', hier=nodes, ranks='genus', maps=True,
             name_as_id=True, note='burst.genus.map: --outmap --name-as-id')
    run_case('blastn_species_map', join('blastn', 'mux.b6o.xz'), hier=nodes,
             ranks='species,none', maps=True,
             note='read maps with taxon:count lists, demultiplexed')
    run_case('blastn_multi_default', join('blastn', 'mux.b6o.xz'), hier=nodes,
             ranks='phylum,genus,species,none,free',
             samples='S01,S03', note='sample whitelist')

    # ordinal: coordinates restricted to the contigs the reads hit
    hit = subjects_of([join(OUT, 'burst', f) for f in
                       sorted(os.listdir(join(OUT, 'burst')))]) | \
        subjects_of([join(OUT, 'bowtie2', f) for f in
                     sorted(os.listdir(join(OUT, 'bowtie2')))][:1])
    if ONLY is None or not os.path.exists(join(OUT, 'coords.txt.xz')):
        with lzma.open(join(fun, 'coords.txt.xz'), 'rt') as fi, \
                lzma.open(join(OUT, 'coords.txt.xz'), 'wt') as fo:
            keep = False
            for line in fi:
                if line[0] in '>#':
                    keep = line[1:].strip() in hit
                if keep:
                    fo.write(line)
    fmaps = dict(map_fps=[join(fun, 'uniref', 'uniref.map.xz'),
                          join(fun, 'go', 'process.tsv.xz')], map_rank=True)
    run_case('burst_orf', 'burst', ranks=None, coords_rel='coords.txt.xz',
             note='coord-match, gene table (cf. bowtie2.orf.tsv)')
    run_case('burst_process', 'burst', hier=fmaps, ranks='process',
             coords_rel='coords.txt.xz', note='burst.process.tsv')
    run_case('burst_genus_process', 'burst', hier=fmaps, ranks='process',
             coords_rel='coords.txt.xz', strata_rel='burst.genus.map',
             note='burst.genus.process.tsv: ordinal + stratified')
    run_case('burst_process_ov55', 'burst', hier=fmaps, ranks='process,none',
             coords_rel='coords.txt.xz', overlap=55,
             note='overlap 55: ceil(len*0.55) differs from exact arithmetic')
    run_case('bowtie2_orf_s01', join('bowtie2', 'S01.sam.xz'), ranks=None,
             coords_rel='coords.txt.xz', overlap=81,
             note='SAM + CIGAR lengths, overlap 81')
    # --sizes (classify.counter_size): a subject-length map for the plain
    # path, gene lengths from the coordinates ('.') for the ordinal path
    os.makedirs(join(OUT, 'sizes'), exist_ok=True)
    subs = sorted(subjects_of([join(OUT, 'bt2sho', f) for f in
                               sorted(os.listdir(join(OUT, 'bt2sho')))]))
    if ONLY is None:
        with open(join(OUT, 'sizes', 'bt2sho.length.map'), 'w') as f:
            for x in subs:
                f.write(f'{x}\t{1000000 + (int(x[1:]) % 977) * 1013}\n')
    run_case('bt2sho_order_sizes', 'bt2sho', hier=nodes,
             ranks='order,genus,none', sizes=join('sizes', 'bt2sho.length.map'),
             note='cf. bt2sho.order.cpm.tsv: size-weighted, multi-hit data')
    run_case('bt2sho_order_sizes_unasgd', 'bt2sho', hier=nodes,
             ranks='order', sizes=join('sizes', 'bt2sho.length.map'),
             unasgd=True, major=60, note='size-weighted, majority, Unassigned')
    run_case('burst_process_sizes', 'burst', hier=fmaps, ranks='process,none',
             coords_rel='coords.txt.xz', sizes='.',
             note='cf. bt2sho.component.rpk.tsv: ordinal, gene lengths as sizes')
    # --sizes together with --stratify (classify.counter_size_strat)
    fp = join(OUT, 'sizes', 'burst.length.map')
    if not os.path.exists(fp):
        bsubs = sorted(subjects_of([join(OUT, 'burst', f) for f in
                                    sorted(os.listdir(join(OUT, 'burst')))]))
        with open(fp, 'w') as f:
            for x in bsubs:
                f.write(f'{x}\t{900000 + (int(x[1:]) % 911) * 1009}\n')
    run_case('burst_ogu_strat_sizes', 'burst', ranks=None,
             strata_rel='burst.genus.map',
             sizes=join('sizes', 'burst.length.map'),
             note='plain: genome counts weighted by length, per genus stratum')
    run_case('burst_ogu_strat_sizes_map', 'burst', ranks=None,
             strata_rel='burst.genus.map', maps=True,
             sizes=join('sizes', 'burst.length.map'),
             note='--sizes + --stratify + --outmap: the map lists every read')
    run_case('burst_genus_process_sizes_map', 'burst', hier=fmaps,
             ranks='process,none', coords_rel='coords.txt.xz',
             strata_rel='burst.genus.map', sizes='.', maps=True,
             note='ordinal + stratified + sizes + read maps')
    run_case('burst_genus_process_sizes', 'burst', hier=fmaps,
             ranks='process,none', coords_rel='coords.txt.xz',
             strata_rel='burst.genus.map', sizes='.',
             note='ordinal + stratified + gene lengths as sizes')
    # subject coverage (--outcov; range.py): SAM with CIGAR spans, multi-hit
    # reads, five samples; BED-like and GFF-like coordinates
    run_case('bowtie2_ogu_cov', 'bowtie2', ranks=None, cov=True,
             note='--outcov: bowtie2 SAM, default (BED-like) coordinates')
    run_case('bt2sho_genus_cov_gff', 'bt2sho', hier=nodes, ranks='genus',
             cov='gff', note='--outcov --outcov-fmt gff with a hierarchy')
    run_case('blastn_species_cov', join('blastn', 'mux.b6o.xz'),
             hier=nodes, ranks='species', cov='1i', samples='S01,S03',
             note='--outcov on demultiplexed b6o, sample whitelist, 1i')
    # read maps of the ordinal path (the gene sets of ordinal_mapper through
    # assign_readmap / write_readmap)
    run_case('burst_process_map', 'burst', hier=fmaps, ranks='process,none',
             coords_rel='coords.txt.xz', maps=True,
             note='--coords with --outmap: gene and process read maps')
    run_case('bowtie2_orf_s01_map', join('bowtie2', 'S01.sam.xz'), ranks=None,
             coords_rel='coords.txt.xz', overlap=81, maps=True,
             note='--coords with --outmap on SAM, no hierarchy')


def synthetic():
    """Modes and edge cases no bundled file has: SAM mates with multi-hits,
    exact duplicate records, subjects missing from the tree, long queries."""
    d = join(OUT, 'synth')
    os.makedirs(d, exist_ok=True)
    tax = synth.Taxonomy(seed=5, level_sizes=[1, 2, 6, 15, 40, 90, 200, 500],
                         n_genomes=1200)
    dt = join(OUT, 'synth_tax')
    os.makedirs(dt, exist_ok=True)
    tax.write_nodes_dmp(join(dt, 'nodes.dmp'))
    tax.write_taxid_map(join(dt, 'taxid.map'))
    rng = np.random.default_rng(77)
    for si in range(3):
        qn, sn, fl = [], [], []
        for qi in range(6000):
            k = min(rng.geometric(0.45), 16)
            if qi % 997 == 0:
                k = 40                       # long query (slow path, d > 16)
            base = int(rng.integers(0, tax.n_genomes))
            paired = rng.random() < 0.3
            for j in range(k):
                g = (base + int(rng.integers(0, 25)) * (j > 0)) % tax.n_genomes
                sid = tax.genome_id(g) if rng.random() > 0.03 else f'X{g}'
                flag = 0
                if paired:
                    flag = 64 if rng.random() < 0.5 else 128
                qn.append(f'S{si}R{qi}')
                sn.append(sid)
                fl.append(flag)
                if rng.random() < 0.04:      # exact duplicate record
                    qn.append(qn[-1])
                    sn.append(sid)
                    fl.append(flag)
        tmp = join(d, f'S{si}.sam')
        synth.write_sam(tmp, qn, sn, flags=fl)
        with open(tmp, 'rb') as fi, lzma.open(tmp + '.xz', 'wb') as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(tmp)
    hier = dict(nodes_fps=[join(dt, 'nodes.dmp')],
                map_fps=[join(dt, 'taxid.map')])
    run_case('synth_default_map', 'synth', hier=hier,
             ranks='phylum,genus,free,none', maps=True,
             note='read maps: mates, duplicates, unknown subjects, 40-hit '
                  'queries')
    run_case('synth_unasgd_map', 'synth', hier=hier, ranks='genus,free',
             maps=True, unasgd=True, uniq=True, note='read maps with Unassigned')
    for mode, kw in [('default', {}), ('uniq', dict(uniq=True)),
                     ('major80', dict(major=80)), ('major54', dict(major=54)),
                     ('major81', dict(major=81)), ('above', dict(above=True)),
                     ('unasgd', dict(unasgd=True)),
                     ('above_unasgd', dict(above=True, unasgd=True))]:
        run_case(f'synth_{mode}', 'synth', hier=hier,
                 ranks='phylum,genus,species,free,none', **kw,
                 note='synthetic SAM: mates, duplicates, unknown subjects, '
                      '40-hit queries')


def synthetic_ordinal():
    """cfg3 at test size through the generator the bench uses (gen_genes,
    gen_reads): reverse-strand genes, multi-hit reads, gapped CIGARs — pins
    the generator's (beg, end, len) columns to what the reference derives
    from the SAM lines."""
    if ONLY is not None and 'synth_ordinal' not in ONLY and \
            'synth_ordinal_ov55' not in ONLY:
        return
    d = join(OUT, 'synth_ordinal')
    os.makedirs(d, exist_ok=True)
    coff, gb, ge = synth.gen_genes(12, 150, 200_000, seed=31)
    synth.write_coords(join(OUT, 'synth_coords.txt'), coff, gb, ge)
    for si in range(2):
        q, c, b, e, ln, _ = synth.gen_reads(4000, 12, 200_000, seed=40 + si)
        assert set((e - b).tolist()) == {150}
        synth.reads_as_sam(join(d, f'S{si}.sam'), q.numpy(), c.numpy(),
                           b.numpy(), ln.numpy())
    run_case('synth_ordinal', 'synth_ordinal', ranks=None,
             coords_rel='synth_coords.txt', maps=True,
             note='cfg3-shaped synthetic reads and genes, overlap 80, maps')
    run_case('synth_ordinal_ov55', 'synth_ordinal', ranks=None,
             coords_rel='synth_coords.txt', overlap=55,
             note='same, overlap 55')


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--only':
        ONLY = set(sys.argv[2].split(','))
        bundled()
        synthetic_ordinal()
        with open(join(HERE, 'INDEX.json')) as f:
            old = json.load(f)
        with open(join(HERE, 'INDEX.json'), 'w') as f:
            json.dump(old + [c for c in CASES if c not in old], f)
        shutil.rmtree(stub)
        sys.exit(0)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    bundled()
    synthetic()
    synthetic_ordinal()
    with open(join(HERE, 'INDEX.json'), 'w') as f:
        json.dump(CASES, f)
    shutil.rmtree(stub)
