"""Golden results of the reference's `build_hierarchy`
(/root/reference/woltka/workflow.py:698-815) on small hierarchy files in the
six formats it reads.  The files are written by `write_inputs` (shared with
tests/test_loaders.py); the reference's (tree, rankdic, namedic, root) for
every case go to tests/golden/hierarchy.json.

    python tests/golden/make_hierarchy_golden.py      (needs /root/reference)
"""
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

FILES = {
    'names.dmp': (
        '1\t|\troot\t|\t\t|\tscientific name\t|\n'
        '1\t|\tall\t|\t\t|\tsynonym\t|\n'
        '2\t|\tBacteria\t|\tBacteria <bacteria>\t|\tscientific name\t|\n'
        '2\t|\teubacteria\t|\t\t|\tgenbank common name\t|\n'
        '10\t|\tAlpha\t|\t\t|\tscientific name\t|\n'
        '11\t|\tAlpha one\t|\t\t|\tscientific name\t|\n'
        '12\t|\tAlpha two\t|\t\t|\tscientific name\t|\n'
        '20\t|\tBeta\t|\t\t|\tscientific name\t|\n'
        '21\t|\tBeta one\t|\t\t|\tscientific name\t|\n'),
    'names.tsv': 'x1\tFirst\nx2\tSecond name\n',
    'nodes.dmp': (
        '1\t|\t1\t|\tno rank\t|\n'
        '2\t|\t1\t|\tsuperkingdom\t|\n'
        '10\t|\t2\t|\tgenus\t|\n'
        '11\t|\t10\t|\tspecies\t|\n'
        '12\t|\t10\t|\tspecies\t|\n'
        '20\t|\t2\t|\tgenus\t|\n'
        '21\t|\t20\t|\tspecies\t|\n'),
    'nodes_norank.tsv': 'a1\tA\na2\tA\nA\tR\nb1\tB\nB\tR\n',
    'nodes_tworoots.tsv': 'a1\tA\tspecies\nb1\tB\tspecies\nA\tA\tgenus\nB\tB\tgenus\n',
    'tree.nwk': "((a1:0.1,'a2':0.2)A:0.5,(b1,b2,(c1,c2)\"C\")B)R;\n",
    'lineages.txt': (
        '# comment\n'
        'G1\tk__Bacteria; p__Firmi; c__; o__Lacto; f__; g__; s__\n'
        'G2\tk__Bacteria; p__Firmi; c__Bacilli; o__Lacto\n'
        'G3\tk__Archaea; p__Eury; Unclassified; x__odd; g__Methano\n'
        'G4\td__Viruses\n'),
    'columns.tsv': (
        '#ID\tphylum\tclass\tgenus\n'
        'G1\tP1\tC1\tGa\n'
        'G2\tP1\tC1\tGb\n'
        'G3\tP1\t\tGc\n'
        'G4\tP2\tunclassified\t0\n'
        'G5\t\t\t\n'),
    'columns_conflict.tsv': '#ID\tphylum\tclass\nG1\tP1\tC1\nG2\tP2\tC1\n',
    'nucl2g.txt': 'n1\tG1\nn2\tG1\textra\nn3\tG2 \nbare\n',
    'g-to-species.map': 'G1\tsp1\nG2\tsp1\nG3\tsp2\n',
    'species2genus.txt': 'sp1\tgen1\nsp2\tgen1\n',
}

CASES = {
    'names_ncbi': dict(names_fps=['names.dmp']),
    'names_plain': dict(names_fps=['names.tsv']),
    'nodes_ncbi': dict(nodes_fps=['nodes.dmp'], names_fps=['names.dmp']),
    'nodes_norank': dict(nodes_fps=['nodes_norank.tsv']),
    'nodes_tworoots': dict(nodes_fps=['nodes_tworoots.tsv']),
    'newick': dict(newick_fps=['tree.nwk']),
    'lineage': dict(lineage_fps=['lineages.txt']),
    'columns': dict(columns_fps=['columns.tsv']),
    'columns_conflict': dict(columns_fps=['columns_conflict.tsv']),
    'map_rank_from_stem': dict(map_fps=['g-to-species.map', 'species2genus.txt']),
    'map_no_rank': dict(map_fps=['nucl2g.txt'], map_rank=False),
    'map_below_tree': dict(newick_fps=['tree.nwk'], map_fps=['nucl2g.txt']),
    'map_below_tree_ranked': dict(nodes_fps=['nodes.dmp'], map_rank=True,
                                  map_fps=['species2genus.txt']),
    'conflict_between_files': dict(nodes_fps=['nodes.dmp', 'nodes_tworoots.tsv'],
                                   map_fps=['g-to-species.map', 'nucl2g.txt']),
    'conflict_nodes': dict(nodes_fps=['nodes_norank.tsv', 'nodes_tworoots.tsv']),
    'nothing': dict(),
}


def write_inputs(dirname):
    for name, text in FILES.items():
        with open(os.path.join(dirname, name), 'w') as f:
            f.write(text)


def run_case(build, dirname, kw):
    """('ok', [tree, rankdic, namedic, root], stdout) or ('err', type, message)."""
    kw = {k: [os.path.join(dirname, x) for x in v] if isinstance(v, list) else v
          for k, v in kw.items()}
    out = io.StringIO()
    try:
        with redirect_stdout(out):
            tree, rankdic, namedic, root = build(**kw)
    except Exception as err:     # what fails in the reference must fail here
        return ['err', type(err).__name__, str(err)]
    return ['ok', [dict(tree), rankdic, namedic, root], out.getvalue()]


if __name__ == '__main__':
    from baseline.reference_arm import find_reference
    wf, where = find_reference()
    assert wf is not None, where
    with tempfile.TemporaryDirectory() as d:
        write_inputs(d)
        golden = {name: run_case(wf.build_hierarchy, d, kw)
                  for name, kw in CASES.items()}
    with open(os.path.join(HERE, 'hierarchy.json'), 'w') as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print({k: v[0] for k, v in golden.items()})
