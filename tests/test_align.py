"""Alignment readers: known answers of the reference's reader tests
(/root/reference/woltka/tests/test_align.py), restated."""
import pytest

from woltka_b200.align import (plain_mapper, iter_align, infer_align_format,
                               cigar_to_lens)

SAM = (
    '@HD	VN:1.0	SO:unsorted',
    'S1	77	NC_123456	26	0	100M	*	0	0	*	*',
    'S1	141	NC_123456	151	0	80M	*	0	0	*	*',
    'S2	0	NC_789012	186	0	50M5I20M5D20M	*	0	0	*	*',
    'S2	16	*	0	0	*	*	0	0	*	*',
    'S3	83	NC_123456	452	0	100M	*	0	0	*	*',
    'S3	163	NC_123456	378	0	80M5D15M	*	0	0	*	*',
    'S3	355	NC_345678	133	0	100M	*	0	0	*	*',
    'S3	403	NC_345678	261	0	10M5I85M	*	0	0	*	*')


def test_sam_mates_and_unmapped():
    # tests/test_align.py:208-237
    obs = list(iter_align(iter(SAM), 'sam'))
    assert obs == [('S1/1', {'NC_123456'}), ('S1/2', {'NC_123456'}),
                   ('S2', {'NC_789012'}),
                   ('S3/1', {'NC_123456', 'NC_345678'}),
                   ('S3/2', {'NC_123456', 'NC_345678'})]
    assert not list(iter_align(iter(SAM[:1]), 'sam'))
    assert not list(iter_align(iter(()), 'sam'))


def test_sam_extra_columns():
    # tests/test_align.py:250-275: (subject, None, length, beg, end)
    obs = list(iter_align(iter(SAM), 'sam', extr=True))
    assert obs == [
        ('S1/1', [('NC_123456', None, 100, 25, 125)]),
        ('S1/2', [('NC_123456', None, 80, 150, 230)]),
        ('S2', [('NC_789012', None, 90, 185, 280)]),
        ('S3/1', [('NC_123456', None, 100, 451, 551),
                  ('NC_345678', None, 100, 132, 232)]),
        ('S3/2', [('NC_123456', None, 95, 377, 477),
                  ('NC_345678', None, 95, 260, 355)])]


def test_sam_exclusion_drops_the_whole_query():
    # tests/test_align.py:277-303
    sam = (
        '@HD	VN:1.0	SO:unsorted',
        'S1	77	G1	81	0	50M	*	0	0	*	*',
        'S1	141	G2	81	0	50M	*	0	0	*	*',
        'S2	0	G2	81	0	50M	*	0	0	*	*',
        'S2	16	G3	81	0	50M	*	0	0	*	*',
        'S2	147	G4	81	0	50M	*	0	0	*	*',
        'S2	99	G3	81	0	50M	*	0	0	*	*',
        'S2	0	*	0	0	*	*	0	0	*	*',
        'S3	83	G3	81	0	50M	*	0	0	*	*',
        'S3	163	G1	81	0	50M	*	0	0	*	*',
        'S3	355	G4	81	0	50M	*	0	0	*	*',
        'S3	403	G5	81	0	50M	*	0	0	*	*',
        'S4	99	G6	81	0	50M	*	0	0	*	*',
        'S4	147	G6	81	0	50M	*	0	0	*	*',
        'S4	256	G6	81	0	50M	*	0	0	*	*')
    obs = list(iter_align(iter(sam), 'sam', excl={'G1'}))
    assert obs == [('S2', {'G2', 'G3'}), ('S2/1', {'G3'}), ('S2/2', {'G4'}),
                   ('S4', {'G6'}), ('S4/1', {'G6'}), ('S4/2', {'G6'})]
    # the last query is dropped too when excluded (the reference's ex_ft
    # variant forgets that, align.py:542-547 — documented deviation)
    obs = list(iter_align(iter(sam), 'sam', excl={'G6'}, extr=True))
    assert [q for q, _ in obs] == ['S1/1', 'S1/2', 'S2', 'S2/1', 'S2/2',
                                   'S3/1', 'S3/2']


def test_cigar():
    # tests/test_align.py:349-355
    assert cigar_to_lens('150M') == (150, 150)
    assert cigar_to_lens('3M1I3M1D5M') == (11, 12)
    assert cigar_to_lens('*') == (0, 0)


def test_b6o_map_paf():
    b6o = ('S1/1	NC_123456	100	100	0	0	1	100	225	324	1.2e-30	345',
           'S1/2	NC_123456	95	98	2	1	2	99	708	608	3.4e-20	270')
    assert list(iter_align(iter(b6o), 'b6o')) == [
        ('S1/1', {'NC_123456'}), ('S1/2', {'NC_123456'})]
    # reversed coordinates are ordered, start is 0-based (align.py:832)
    assert list(iter_align(iter(b6o), 'b6o', extr=True)) == [
        ('S1/1', [('NC_123456', 345.0, 100, 224, 324)]),
        ('S1/2', [('NC_123456', 270.0, 98, 607, 708)])]
    tsv = ('R1	A', 'R1	B	x', 'bad line', 'R2	C', 'R1	A')
    assert list(iter_align(iter(tsv), 'map')) == [
        ('R1', {'A', 'B'}), ('R2', {'C'}), ('R1', {'A'})]
    assert list(iter_align(iter(tsv), 'map', excl={'B'})) == [
        ('R2', {'C'}), ('R1', {'A'})]
    paf = ('q1	150	0	150	+	G1	5000	100	250	150	150	60',
           'q1	150	0	150	-	G2	5000	300	445	140	145	60',
           'too	short')
    assert list(iter_align(iter(paf), 'paf')) == [('q1', {'G1', 'G2'})]
    assert list(iter_align(iter(paf), 'paf', extr=True)) == [
        ('q1', [('G1', 60, 150, 100, 250), ('G2', 60, 145, 300, 445)])]


def test_format_inference_and_errors():
    assert infer_align_format(iter(SAM))[0] == 'sam'
    assert infer_align_format(iter(SAM[1:]))[0] == 'sam'
    assert infer_align_format(iter(('R1	A',)))[0] == 'map'
    b6o = 'S1	G1	100	100	0	0	1	100	225	324	1.2e-30	345'
    assert infer_align_format(iter((b6o,)))[0] == 'b6o'
    paf = 'q1	150	0	150	+	G1	5000	100	250	150	150	60'
    assert infer_align_format(iter((paf,)))[0] == 'paf'
    with pytest.raises(ValueError, match='empty or unreadable'):
        infer_align_format(iter(()))
    with pytest.raises(ValueError, match='Cannot determine'):
        infer_align_format(iter(('just text',)))
    with pytest.raises(ValueError, match='Invalid format code'):
        list(iter_align(iter(SAM), 'xyz'))


def test_plain_mapper_chunks():
    # tests/test_align.py:35-115: n queries per chunk, last one short
    lines = [f'R{i}	G{i % 3}' for i in range(7)]
    for n in (1, 2, 3, 5, 7, 10):
        chunks = list(plain_mapper(iter(lines), fmt='map', n=n))
        assert [len(q) for q, _ in chunks] == \
            [n] * (7 // n) + ([7 % n] if 7 % n else [])
        assert sum((q for q, _ in chunks), []) == [f'R{i}' for i in range(7)]
    assert plain_mapper.__name__ == 'plain_mapper'


def test_generator_columns_equal_what_the_readers_derive_from_its_sam_lines():
    """The cfg3 bench feeds gen_reads' integer columns; the golden case
    `synth_ordinal` is the same reads as SAM text through the real reference.
    The two meet here: the readers turn the SAM lines back into exactly the
    generator's (contig, beg, end, aligned length) columns."""
    from os.path import join, dirname, abspath
    import numpy as np
    from woltka_b200 import synth
    from woltka_b200.ordinal import iter_records
    data = join(dirname(abspath(__file__)), 'golden', 'data', 'synth_ordinal')
    for si in range(2):
        q, c, b, e, ln, _ = synth.gen_reads(4000, 12, 200_000, seed=40 + si)
        with open(join(data, f'S{si}.sam')) as fh:
            qn, cn, bg, en, le = next(iter_records(iter(fh), 'sam'))
        assert qn == [f'R{x}' for x in q.tolist()]
        assert cn == [f'C{x}' for x in c.tolist()]
        assert np.array_equal(bg, b.numpy()) and np.array_equal(en, e.numpy())
        assert np.array_equal(le, ln.numpy())
        assert set(le) == {148, 150}


def test_demultiplex_with_an_empty_sample_list_keeps_everything():
    """workflow.demultiplex (workflow.py:885-886) tests `if samples`: an empty
    list is no filter."""
    from woltka_b200.workflow import demultiplex
    qs, ss = ['S1_a', 'S2_b', 'c'], [{'x'}, {'y'}, {'z'}]
    assert demultiplex(qs, ss, []) == demultiplex(qs, ss, None) == {
        'S1': (['a'], [{'x'}]), 'S2': (['b'], [{'y'}]), '': (['c'], [{'z'}])}
    assert demultiplex(qs, ss, ['S2']) == {'S2': (['b'], [{'y'}])}
