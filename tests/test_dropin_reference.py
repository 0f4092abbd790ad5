"""The drop-in claim, proven through the reference's own entry point.

When the unmodified reference is importable on this box (/root/reference in the
build container, or a driver-side install under baseline/_ref), its
`woltka.workflow.classify` and `build_mapper` are replaced by
`woltka_b200.workflow.classify` / `build_mapper` and the reference's own
`woltka classify` command (cli.classify_cmd -> workflow.workflow, cli.py:195-199,
workflow.py:136-141) is run for every command of its CLI test
(woltka/tests/test_cli.py:42-177).  The tables it writes must be byte-identical
to the reference's golden outputs (tests/data/output/*.tsv), the read maps
line-identical.  Everything but the two seams is the reference's code: sample
discovery, hierarchy readers, frac / scale / round, the TSV writer.

CPU: the oracle stands in for the GPU behind the engine interface; `-m gpu`:
the CUDA engine through the C-ABI.  Skipped where there is no reference.
"""
import gzip
import os
import sys
from filecmp import cmp
from os.path import join

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import reference_arm          # noqa: E402

wf, WHERE = reference_arm.find_reference()
pytestmark = pytest.mark.skipif(wf is None, reason=f'no reference: {WHERE}')


def commands(datdir, output_fp, tmpdir):
    aln, tax, fun, out = (join(datdir, x) for x in
                          ('align', 'taxonomy', 'function', 'output'))
    return [
        ('bowtie2.ogu.tsv', ['--input', join(aln, 'bowtie2')]),
        ('bowtie2.free.tsv', ['--input', join(aln, 'bowtie2'),
                              '--nodes', join(tax, 'nodes.dmp'),
                              '--map', join(tax, 'taxid.map'),
                              '--rank', 'free']),
        ('blastn.species.tsv', ['--input', join(aln, 'blastn', 'mux.b6o.xz'),
                                '--lineage', join(tax, 'lineages.txt'),
                                '--rank', 'species']),
        ('burst.genus.tsv', ['--input', join(aln, 'burst'),
                             '--outmap', tmpdir,
                             '--names', join(tax, 'names.dmp'),
                             '--nodes', join(tax, 'nodes.dmp'),
                             '--map', join(tax, 'taxid.map'),
                             '--rank', 'genus', '--name-as-id']),
        ('blastn.family.percent.tsv', [
            '--input', join(aln, 'blastn', 'mux.b6o.xz'),
            '--names', join(tax, 'names.dmp'), '--nodes', join(tax, 'nodes.dmp'),
            '--map', join(tax, 'taxid.map'), '--rank', 'family',
            '--name-as-id', '--frac', '--scale', '100', '--digits', 2]),
        ('bt2sho.order.cpm.tsv', [
            '--input', join(aln, 'bt2sho'), '--names', join(tax, 'names.dmp'),
            '--nodes', join(tax, 'nodes.dmp'), '--map', join(tax, 'taxid.map'),
            '--rank', 'order', '--sizes', join(tax, 'length.map'),
            '--scale', '1M', '--digits', 3]),
        ('bt2sho.phylo.tsv', ['--input', join(aln, 'bt2sho'),
                              '--newick', join(datdir, 'tree.nwk'),
                              '--rank', 'free', '--subok']),
        ('burst.process.tsv', [
            '--input', join(aln, 'burst'), '--rank', 'process',
            '--coords', join(fun, 'coords.txt.xz'),
            '--map', join(fun, 'uniref', 'uniref.map.xz'),
            '--map', join(fun, 'go', 'process.tsv.xz')]),
        ('burst.genus.process.tsv', [
            '--input', join(aln, 'burst'), '--rank', 'process',
            '--coords', join(fun, 'coords.txt.xz'),
            '--map', join(fun, 'uniref', 'uniref.map.xz'),
            '--map', join(fun, 'go', 'process.tsv.xz'),
            '--stratify', join(out, 'burst.genus.map')]),
        ('bt2sho.component.rpk.tsv', [
            '--input', join(aln, 'bt2sho'), '--rank', 'component',
            '--coords', join(fun, 'coords.txt.xz'),
            '--map', join(fun, 'uniref', 'uniref.map.xz'),
            '--map', join(fun, 'go', 'component.tsv.xz'),
            '--sizes', '.', '--scale', '1k', '--digits', 3]),
        ('split.genus.tsv', [
            '--input', join(aln, 'burst', 'split'), '--trim-sub', '_',
            '--rank', 'genus', '--map', join(tax, 'nucl', 'nucl2tid.txt'),
            '--names', join(tax, 'names.dmp'), '--nodes', join(tax, 'nodes.dmp'),
            '--name-as-id']),
        ('split.process.tsv', [
            '--input', join(aln, 'burst', 'split'), '--rank', 'process',
            '--map', join(fun, 'nucl', 'uniref.map.xz'),
            '--map', join(fun, 'go', 'process.tsv.xz')]),
    ]


def run_all(tmp_path, monkeypatch, engine):
    from click.testing import CliRunner
    from woltka.cli import classify_cmd
    from woltka_b200 import workflow as ours
    from tests.oracle_engine import make_factory

    import inspect
    sig = inspect.signature(ours.classify)

    def classify(*args, **kwargs):
        # (the reference passes all 26 arguments by position, workflow.py:138-141)
        kw = sig.bind(*args, **kwargs).arguments
        if engine == 'oracle':
            kw['_engine_factory'] = make_factory(
                kw.get('tree'), kw.get('rankdic'), kw.get('root'),
                kw.get('ranks'), kw.get('subok', False))
        return ours.classify(**kw)

    # the seams (workflow.py:86, :128 and :138)
    monkeypatch.setattr(wf, 'classify', classify)
    monkeypatch.setattr(wf, 'build_mapper', ours.build_mapper)
    monkeypatch.setattr(wf, 'build_hierarchy', ours.build_hierarchy)
    datdir = join(WHERE, 'woltka', 'tests', 'data')
    output_fp = str(tmp_path / 'output.tsv')
    runner = CliRunner()
    done = []
    for exp, params in commands(datdir, output_fp, str(tmp_path)):
        res = runner.invoke(classify_cmd, [str(x) for x in params] +
                            ['--output', output_fp, '--no-exe'])
        assert res.exit_code == 0, (exp, res.output, res.exception)
        assert cmp(output_fp, join(datdir, 'output', exp), shallow=False), exp
        if exp == 'burst.genus.tsv':            # its read maps, line by line
            for i in range(1, 6):
                fp = str(tmp_path / f'S0{i}.txt.gz')
                with gzip.open(fp, 'rt') as f:
                    obs = [x.rstrip() for x in f]
                with gzip.open(join(datdir, 'output', 'burst.genus.map',
                                    f'S0{i}.txt.gz'), 'rt') as f:
                    assert obs == [x.rstrip() for x in f], i
                os.remove(fp)
        done.append(exp)
    assert len(done) == 12


def test_reference_workflow_through_the_drop_in_oracle(tmp_path, monkeypatch):
    run_all(tmp_path, monkeypatch, 'oracle')


@pytest.mark.gpu
def test_reference_workflow_through_the_drop_in_gpu(tmp_path, monkeypatch):
    run_all(tmp_path, monkeypatch, 'gpu')
