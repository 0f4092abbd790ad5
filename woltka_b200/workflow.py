"""Drop-in replacements for the two seams of the reference's classify path:

    build_mapper(coords_fp, outcov_dir, overlap, chunk, zippers)
    classify(mapper, files, samples, fmt, demux, trimsub, tree, rankdic,
             namedic, root, ranks, rank2dir, outzip, uniq, major, above,
             subok, sizes, unasgd, stratmap, exclude, chunk, cache, zippers,
             outcov_dir, outcov_fmt) -> {rank: {sample: {feature: count}}}

(/root/reference/woltka/workflow.py:536-585 and :162-353; same names,
argument meaning, defaults and error texts.)  The per-chunk body of the
reference — demultiplex, strip_suffix, one assigner + counter per rank,
sum_dict — runs on the GPU through woltka_b200.session / the C-ABI; there is
no CPU fallback.  Everything around it (sample discovery, hierarchy readers,
frac/scale/round, table writers) is the reference's own code and is called
unchanged by its `workflow()`.

Read maps (`rank2dir`) come from a per-record
assignment column the kernel writes next to the counts; size-weighted counts
(`sizes`, row F4) from the kernels' exact (subject, feature) shares, weighted
on the host.  With `--coords` the read maps are made from the matcher's
(record, gene) pairs sent through the plain path, in the reference's line
order.  Subject coverage (`outcov_dir`, row F5) is accumulated and merged on
the GPU (woltka_b200.coverage).  With `sizes` and
`stratmap` together (classify.counter_size_strat) a subject seen in a stratum
becomes its own device subject, so the same (subject, feature) shares also
carry the stratum.
"""
import bz2
import gzip
import lzma
import os
import sys
from functools import partial
from os.path import basename

from .align import plain_mapper
from .ordinal import (load_gene_coords, load_gene_coords_cached,
                      ordinal_mapper, GeneIndex, iter_records)
from .session import Session, _split_sample
from ._lib import WoltkaB200Error
from .coverage import range_mapper, Coverage, coverage_offsets
from .reader import BlockReader
from .loaders import build_hierarchy

__all__ = ['classify', 'build_mapper', 'build_hierarchy', 'assign_readmap', 'demultiplex',
           'strip_suffix', 'read_strata', 'readzip', 'range_mapper']

_OPENERS = {'.gz': gzip.open, '.bz2': bz2.open, '.xz': lzma.open,
            '.lzma': lzma.open}


def readzip(fp, zippers=None):
    """Text handle on a plain or gz/bz2/xz-compressed file (file.py:62-129;
    `zippers` is accepted for signature compatibility — decompression is done
    in-process)."""
    for ext, opener in _OPENERS.items():
        if fp.endswith(ext):
            return opener(fp, 'rt')
    return open(fp, 'r')


def openzip(fp, mode='rt'):
    """Open a plain or gz/bz2/xz file by extension (file.py:31-59)."""
    for ext, opener in _OPENERS.items():
        if fp.endswith(ext):
            return opener(fp, mode)
    return open(fp, mode.replace('t', '') if 'b' in mode else mode)


def _host_cut(text):
    """Bytes of a block that end at a line end and where the query name
    changes (a query is never split, align.py:73-79): what wk_parse_block
    consumes, for the host reader."""
    end = text.rfind(b'\n')
    if end < 0:
        return 0
    # back over the trailing lines that share the last query name
    ls = text.rfind(b'\n', 0, end) + 1
    name = text[ls:end].split(b'\t', 1)[0]
    cut = ls
    while cut > 0:
        ps = text.rfind(b'\n', 0, cut - 1) + 1
        if text[ps:cut - 1].split(b'\t', 1)[0] != name:
            break
        cut = ps
    return cut


# which reader fed the last file of the last classify() call ('device' | 'host')
LAST_READER = None


def _echo(msg, nl=True):
    sys.stdout.write(msg + ('\n' if nl else ''))
    sys.stdout.flush()


def build_mapper(coords_fp=None, outcov_dir=None, overlap=None, chunk=None,
                 zippers=None):
    """Plain or ordinal mapper and its chunk size (workflow.py:536-585)."""
    if coords_fp:
        _echo('Reading gene coordinates...', nl=False)
        cache = os.environ.get('WOLTKA_B200_CACHE')
        if cache:
            # binary cache of the parsed table
            coords, idmap, prefix = load_gene_coords_cached(
                coords_fp, partial(readzip, zippers=zippers), cache)
        else:
            with readzip(coords_fp, zippers) as fh:
                coords, idmap, prefix = load_gene_coords(fh, sort=True)
        _echo(' Done.')
        _echo(f'  Total number of host sequences: {len(coords)}.')
        chunk = chunk or 2 ** 20
        return partial(ordinal_mapper, coords=coords, idmap=idmap,
                       prefix=prefix, th=overlap and overlap / 100), chunk
    return (range_mapper if outcov_dir else plain_mapper), chunk or 1024


def _device_format(fileobj, fmt):
    """Format of the file if the device reader knows it: given, or inferred
    from its first line like align.infer_align_format (align.py:153-223)."""
    if not fmt:
        from .align import infer_align_format
        try:
            pos = fileobj.tell()
            fmt, _ = infer_align_format(fileobj)
            fileobj.seek(pos)
        except (ValueError, OSError):
            return None
    return fmt if fmt in ('sam', 'b6o', 'paf', 'map') else None


def _read_strata(fp, zippers=None):
    """Read-to-stratum map of one sample (workflow.py:912-938 with
    file.read_map_uniq, file.py:368-385: only two-column lines count)."""
    strata = {}
    with readzip(fp, zippers) as fh:
        for line in fh:
            key, found, value = line.partition('\t')
            if found and '\t' not in value:
                strata[key] = value.rstrip()
    if not strata:
        raise ValueError('No stratification information is found in file: '
                         f'{basename(fp)}.')
    return strata


def classify(mapper, files, samples=None, fmt=None, demux=None, trimsub=None,
             tree=None, rankdic=None, namedic=None, root=None, ranks=None,
             rank2dir=None, outzip=None, uniq=False, major=None, above=False,
             subok=False, sizes=None, unasgd=False, stratmap=None,
             exclude=None, chunk=None, cache=1024, zippers=None,
             outcov_dir=None, outcov_fmt=None, _engine_factory=None,
             _device=None):
    """Core of the classification workflow (workflow.py:162-353) on the GPU.

    Under `torchrun` (torch.distributed initialised, one process per GPU) the
    alignment files are dealt out to the ranks, every rank classifies its
    share on its own GPU and the exact per-rank profiles are summed on rank 0
    — the reference's "split, run, `woltka merge`" recipe
    (/root/reference/doc/perform.md:70-98) inside one call.  Rank 0 returns
    the merged profiles, the other ranks empty ones (`is_output_rank()` tells
    a caller which process writes the tables).

    `cache` (LRU size of the reference's assigners) has no effect on results
    and is ignored.  Counts are exact: integers where the reference holds
    integers, and the correctly rounded value of the exact rational sum where
    the reference accumulates 1/k shares in floating point; after
    `round_profiles` the two are identical.
    """
    is_ordinal = getattr(mapper, 'func', None) is ordinal_mapper
    world, rank = _dist_info()
    if _device is None:
        _device = int(os.environ.get('LOCAL_RANK', '0')) if world > 1 else 0
    # side files are appended per sample: share the work only when no two
    # ranks can meet the same sample (one file per sample)
    if world > 1 and ('-' in files or (demux and (rank2dir or outcov_dir))):
        world = 1
        if rank != 0:
            files = type(files)()
    if outcov_dir:
        coverage_offsets(outcov_fmt)     # an invalid format fails up front
        if is_ordinal:
            # (the reference breaks on the gene sets of ordinal_mapper,
            # range.py:142: `subjects.items()`)
            raise ValueError('Subject coverage (--outcov) cannot be combined '
                             'with --coords.')

    genes = None
    if is_ordinal:
        kw = mapper.keywords
        genes = GeneIndex(kw['coords'], kw['idmap'], kw.get('prefix', False))
        th = kw.get('th', 0.8)

    sess = Session(ranks, tree, rankdic, root, uniq, major and major / 100,
                   above, subok, unasgd, trimsub, _engine_factory, _device,
                   rank2dir, outzip, namedic, sizes, bool(stratmap))
    samset = set(samples) if samples else None
    strata_cache = {}
    cover = Coverage(sess.engines[0]) if outcov_dir else None

    def strata_of(sname):
        try:
            return strata_cache[sname]
        except KeyError:
            st = strata_cache[sname] = _read_strata(stratmap[sname], zippers)
            return st

    file_order = {fp: i for i, fp in enumerate(sorted(files))}
    try:
        for fp in sorted(files)[rank::world] if world > 1 else sorted(files):
            sess.file_index = file_index = file_order[fp]
            if fp == '-':
                fileobj = sys.stdin
                _echo('Parsing alignment from stdin ', nl=False)
            else:
                fileobj = readzip(fp, zippers)
                _echo(f'Parsing alignment file {basename(fp)} ', nl=False)
            sname = None if demux else (files[fp] if isinstance(files, dict)
                                        else None)
            kwargs = dict(demux=bool(demux), sample_name=sname,
                          samples=samset if demux else None,
                          strata_of=strata_of if stratmap else None)
            nqry, nstep = 0, -1
            on_device = (fp != '-' and not stratmap and
                         (is_ordinal or mapper is plain_mapper) and
                         not (is_ordinal and rank2dir) and
                         not os.environ.get('WOLTKA_B200_HOST_READER') and
                         sess.can_parse_on_device())
            dfmt = _device_format(fileobj, fmt) if on_device else None
            on_device = dfmt is not None and not (is_ordinal and dfmt == 'map')
            global LAST_READER
            LAST_READER = 'device' if on_device else 'host'
            try:
                if on_device:
                    # text -> columns (-> read-gene matches) -> counts, all on
                    # the device; the host moves blocks of bytes
                    sess.configure_reader(exclude, coords=is_ordinal)
                    common = (bool(demux), sname, samset if demux else None,
                              dfmt)

                    def host_chunk(text):
                        lines = iter(text.decode().splitlines(True))
                        if not is_ordinal:
                            return sess.add_text_chunk_host(
                                text, *common, chunk or 1024, exclude)
                        n = 0
                        for qn, cn, bg, en, ln in iter_records(
                                lines, dfmt, exclude, chunk or 2 ** 20):
                            n += len(set(qn))
                            sess.add_ordinal_chunk(genes, qn, cn, bg, en, ln,
                                                   th, **kwargs)
                        return n

                    blocks = BlockReader(fp, header=dfmt == 'sam')
                    for view, final in blocks:
                        if sess.can_parse_on_device():
                            try:
                                if is_ordinal:
                                    used, n = sess.add_text_block_ordinal(
                                        view, final, genes, th, *common)
                                else:
                                    used, n = sess.add_text_block(
                                        view, final, *common)
                                blocks.consumed(used)
                                nqry += n
                                continue
                            except WoltkaB200Error as err:
                                if err.code not in (5, 6):
                                    raise
                                LAST_READER = 'host'
                                if err.code == 5:      # WK_ERR_CAPACITY
                                    # the device reader's tables are full (> 1M
                                    # subjects, 64 MiB of names, 32k samples
                                    # or a 64k-line query): this block and
                                    # the rest of the run go through the host
                                    # reader
                                    sess.device_reader_off = True
                                # (6 = WK_ERR_FALLBACK: a query name in two
                                # places of this block, which the reference
                                # merges, ordinal.py:332 - this block only)
                        text = view.tobytes()
                        used = len(text) if final else _host_cut(text)
                        blocks.consumed(used)
                        if used:
                            nqry += host_chunk(text[:used])
                        istep = nqry // 1000000 - nstep
                        if istep:
                            _echo('.' * istep, nl=False)
                            nstep += istep
                elif is_ordinal:
                    for qn, cn, bg, en, ln in iter_records(
                            iter(fileobj), fmt, exclude, chunk or 2 ** 20):
                        nqry += len(set(qn))
                        sess.add_ordinal_chunk(genes, qn, cn, bg, en, ln, th,
                                               **kwargs)
                else:
                    for qryque, subque in mapper(iter(fileobj), fmt=fmt,
                                                 excl=exclude, n=chunk):
                        nqry += len(qryque)
                        if cover is not None:
                            # range.parse_ranges on the demultiplexed chunk
                            # (workflow.py:312-313)
                            for query, ranges in zip(qryque, subque):
                                sam = _split_sample(query)[0] if demux \
                                    else sname
                                if not demux or samset is None or \
                                        sam in samset:
                                    cover.add(sam, ranges)
                        sess.add_chunk(qryque, subque, **kwargs)
                        istep = nqry // 1000000 - nstep
                        if istep:
                            _echo('.' * istep, nl=False)
                            nstep += istep
            finally:
                if fileobj is not sys.stdin:
                    fileobj.close()
            _echo(' Done.')
            _echo(f'  Number of sequences classified: {nqry}.')
        if cover is not None:
            _echo('Calculating per sample coverage...', nl=False)
            cover.write(outcov_dir, outcov_fmt)
            _echo(' Done.')
        _echo('Classification completed.')
        if _dist_info()[0] > 1:
            data = _merge_ranks(sess)
        else:
            data = sess.results()
    finally:
        sess.close()
    # one (possibly empty) profile per requested rank, like workflow.py:268
    return {rank: data[rank] for rank in dict.fromkeys(ranks)}


def _dist_info():
    """(world size, rank) of torch.distributed when it is initialised."""
    if 'torch' not in sys.modules:
        return 1, 0
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def is_output_rank():
    """True in the process that holds the merged profiles of `classify()`."""
    return _dist_info()[1] == 0


def _merge_ranks(sess):
    """Sum the exact per-rank profiles on rank 0 (`woltka merge`,
    tools.py:153-208); samples come back in the order a single process meets
    them (sorted files, first appearance)."""
    from .distributed import merge_profiles
    from .session import finalize
    import torch.distributed as dist
    exact = sess.exact_results()
    seen = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(sess.sample_seen, seen, dst=0)
    merged = merge_profiles(exact)
    if merged is None:
        return {rk: {} for rk in sess.order}
    first = {}
    for part in seen:
        for name, key in part.items():
            if name not in first or key < first[name]:
                first[name] = key
    out = {}
    for rk in sess.order:
        row = merged.get(rk, {})
        out[rk] = {name: row[name] for name in sorted(row, key=first.get)}
    return finalize(out)


def assign_readmap(qryque, subque, data, rank, sample, assigners, cache=1024,
                   rank2dir=None, outzip=None, tree=None, rankdic=None,
                   namedic=None, root=None, uniq=False, major=None,
                   above=False, subok=False, sizes=None, unasgd=False,
                   strata=None, _engine_factory=None, _device=0):
    """The per-chunk seam of the reference (workflow.py:941-1058): assign the
    queries of one (qryque, subque) chunk of `sample` at `rank` on the GPU,
    optionally append the read map, and add the counts into
    `data[rank][sample]` (util.sum_dict).  `major` is the fraction here (the
    reference divides the percentage before it calls this, workflow.py:276);
    `assigners` and `cache` (the reference's memoised assigner closures) are
    accepted and not needed."""
    sess = Session([rank], tree, rankdic, root, uniq, major, above, subok,
                   unasgd, None, _engine_factory, _device, rank2dir, outzip,
                   namedic, sizes, strata is not None)
    try:
        sess.add_chunk(qryque, subque, demux=False, sample_name=sample,
                       strata_of=(lambda _: strata) if strata is not None
                       else None)
        counts = sess.results()[rank].get(sample, {})
    finally:
        sess.close()
    total = data[rank].setdefault(sample, {})
    for key, value in counts.items():
        total[key] = total.get(key, 0) + value


# The three host-side helpers of the chunk loop under their reference names
# (the GPU path does the same work inside Session.add_chunk while it interns
# the strings; these are for callers and tests that use them directly).

def strip_suffix(subque, sep):
    """Subject sets with everything from the last `sep` on removed
    (workflow.py:818-841); trimmed names that coincide fall together."""
    return ({name.rsplit(sep, 1)[0] for name in subjects}
            for subjects in subque)


def demultiplex(qryque, subque, samples=None, sep='_'):
    """{sample: (reads, subjects)} of a multiplexed chunk (workflow.py:844-909):
    the sample is the text before the first `sep` when something follows it,
    else ''; with `samples` only those are kept.  Samples and reads keep the
    order of the chunk."""
    keep = set(samples) if samples else None
    res = {}
    for query, subjects in zip(qryque, subque):
        left, _, right = query.partition(sep)
        sample, read = (left, right) if right else ('', right or left)
        if keep is None or sample in keep:
            reads, subs = res.setdefault(sample, ([], []))
            reads.append(read)
            subs.append(subjects)
    return res


def read_strata(fp, zippers=None):
    """Read-to-stratum map of one sample (workflow.py:912-938)."""
    return _read_strata(fp, zippers)
