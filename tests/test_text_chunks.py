"""Host side of the device SAM reader: the file is cut into byte chunks that
end where the query name changes (no GPU needed)."""
from tests.test_gpu_parse import synthetic_sam


def test_text_chunks_never_split_a_query(tmp_path):
    from woltka_b200 import workflow
    body = synthetic_sam(2000, 9)
    fp = tmp_path / 'x.sam'
    fp.write_bytes(b'@HD\tVN:1.0\n' + body)
    chunks = list(workflow._text_chunks(str(fp), block=7000))
    assert b''.join(chunks) == body and len(chunks) > 10
    for a, b in zip(chunks[:-1], chunks[1:]):
        assert a.endswith(b'\n')
        last = a.rstrip(b'\n').rsplit(b'\n', 1)[-1].split(b'\t', 1)[0]
        assert b.split(b'\t', 1)[0] != last


def test_text_chunks_compressed_and_headers(tmp_path):
    import gzip
    from woltka_b200 import workflow
    body = synthetic_sam(300, 4, trailing_newline=False)
    fp = tmp_path / 'x.sam.gz'
    with gzip.open(fp, 'wb') as f:
        f.write(b'@HD\tVN:1.0\n@PG\tID:x\n' + body)
    assert b''.join(workflow._text_chunks(str(fp), block=5000)) == body
    empty = tmp_path / 'e.sam'
    empty.write_bytes(b'@HD\tVN:1.0\n')
    assert list(workflow._text_chunks(str(empty))) == []
