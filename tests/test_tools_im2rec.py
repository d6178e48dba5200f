"""tools/im2rec.py (list creation, splits, threaded packing with resize / crop) and the native pass-through packer tools/im2rec.cc; both
produce databases that ``mx.recordio`` / ``mx.image.ImageIter`` read."""
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _images(root):
    from PIL import Image
    rng = np.random.RandomState(0)
    for c, cls in enumerate(["cat", "dog", "eel"]):
        os.makedirs(os.path.join(root, cls))
        for i in range(4):
            Image.fromarray(rng.randint(0, 255, (20 + 4 * i, 30, 3), dtype=np.uint8)).save(os.path.join(root, cls, "%d.png" % i))
    open(os.path.join(root, "cat", "notes.txt"), "w").write("not an image")


def test_im2rec_py_list_and_pack(tmp_path):
    import im2rec
    import geomx_b200 as mx
    root = str(tmp_path / "img"); _images(root)
    prefix = str(tmp_path / "db" / "set"); os.makedirs(os.path.dirname(prefix))
    assert im2rec.main(["--list", "--recursive", "--train-ratio", "0.75", prefix, root]) == 0
    train = list(im2rec.read_list(prefix + "_train.lst")); val = list(im2rec.read_list(prefix + "_val.lst"))
    assert len(train) == 9 and len(val) == 3 and not os.path.exists(prefix + "_test.lst")
    assert sorted(i for i, _, _ in train + val) == list(range(12))
    assert {os.path.dirname(p): l[0] for _, l, p in train + val} == {"cat": 0.0, "dog": 1.0, "eel": 2.0}
    assert im2rec.main(["--resize", "16", "--center-crop", "--num-thread", "3", "--encoding", ".png", prefix, root]) == 0      # packs both lists
    rec = mx.recordio.MXIndexedRecordIO(prefix + "_train.idx", prefix + "_train.rec", "r")
    assert sorted(rec.keys) == sorted(i for i, _, _ in train)
    for i, labels, _ in train:
        header, img = mx.recordio.unpack_img(rec.read_idx(i))
        assert header.label == labels[0] and header.id == i and img.shape == (16, 16, 3)
    one = mx.recordio.MXRecordIO(prefix + "_train.rec", "r").read()
    assert mx.recordio.unpack(one)[0].id == train[0][0]                 # records are in list order whatever the thread count
    # chunks, flat listing, no shuffle
    flat = str(tmp_path / "flat"); os.makedirs(flat)
    for i in range(5):
        shutil.copy(os.path.join(root, "cat", "0.png"), os.path.join(flat, "%d.png" % i))
    p2 = str(tmp_path / "db" / "flat")
    assert im2rec.main(["--list", "--no-shuffle", "--chunks", "2", p2, flat]) == 0
    assert [p for _, _, p in im2rec.read_list(p2 + "_0.lst")] == ["0.png", "1.png", "2.png"] and len(list(im2rec.read_list(p2 + "_1.lst"))) == 2
    # multi-label lists are packed as label vectors; a broken file is skipped, not fatal
    p3 = str(tmp_path / "db" / "det")
    with open(p3 + ".lst", "w") as f:
        f.write("0\t2\t5\t0.1\t0.2\t0.3\tcat/0.png\n1\t1.0\tcat/notes.txt\n")
    assert im2rec.main(["--pack-label", p3, root]) == 0
    r = mx.recordio.MXIndexedRecordIO(p3 + ".idx", p3 + ".rec", "r")
    h, _ = mx.recordio.unpack(r.read_idx(0))
    assert r.keys == [0] and np.allclose(h.label, [2, 5, 0.1, 0.2, 0.3])
    assert im2rec.main([str(tmp_path / "nothing"), root]) == 1


def test_native_im2rec_matches_python_pass_through(tmp_path):
    import im2rec
    import geomx_b200 as mx
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "im2rec")
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Werror", "-pthread", os.path.join(ROOT, "tools", "im2rec.cc"), "-o", exe], check=True)
    root = str(tmp_path / "img"); _images(root)
    magic = struct.pack("<I", 0xced7230a)
    open(os.path.join(root, "eel", "magic.png"), "wb").write(b"abcd" + magic + b"payload with the framing word inside" + magic)
    prefix = str(tmp_path / "set")
    assert im2rec.main(["--list", "--recursive", prefix, root]) == 0
    with open(prefix + ".lst", "a") as f:
        f.write("99\t1.000000\tmissing/none.png\n")
    shutil.copy(prefix + ".lst", prefix + "_native.lst")
    r = subprocess.run([exe, prefix + "_native.lst", root, prefix + "_native.rec", "threads=3"], capture_output=True, text=True)
    assert r.returncode == 0 and "packed 13 of 14" in r.stdout and "cannot read" in r.stderr, r.stdout + r.stderr
    os.remove(prefix + "_native.lst")
    assert im2rec.main(["--pass-through", prefix, root]) == 0
    assert open(prefix + ".rec", "rb").read() == open(prefix + "_native.rec", "rb").read()              # byte-identical databases
    assert open(prefix + ".idx").read() == open(prefix + "_native.idx").read()
    rec = mx.recordio.MXIndexedRecordIO(prefix + "_native.idx", prefix + "_native.rec", "r")
    items = {i: (l, p) for i, l, p in im2rec.read_list(prefix + ".lst")}
    for k in rec.keys:
        h, payload = mx.recordio.unpack(rec.read_idx(k))
        assert h.id == k and h.label == items[k][0][0] and payload == open(os.path.join(root, items[k][1]), "rb").read()
    # partitions + packed labels
    r = subprocess.run([exe, prefix + ".lst", root, str(tmp_path / "p.rec"), "nsplit=2", "part=1", "pack_label=1"], capture_output=True, text=True)
    assert r.returncode == 0 and "of 7 images" in r.stdout
    h, _ = mx.recordio.unpack(mx.recordio.MXRecordIO(str(tmp_path / "p.rec.part1"), "r").read())
    assert h.flag == 1 and len(np.atleast_1d(h.label)) == 1
    assert subprocess.run([exe, prefix + ".lst", root, str(tmp_path / "q.rec"), "bogus=1"], capture_output=True).returncode == 1
