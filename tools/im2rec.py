#!/usr/bin/env python
"""Image folder -> ``.lst`` -> ``.rec`` + ``.idx`` (the data format of ``ImageRecordIter`` / ``ImageIter``).

Same command line as the reference's ``tools/im2rec.py`` (``prefix root`` + the list options ``--list --exts --chunks --train-ratio --test-ratio
--recursive --no-shuffle`` and the packing options ``--pass-through --resize --center-crop --quality --num-thread --color --encoding
--pack-label``), built differently: one ``Lister`` that enumerates / splits, one ``Packer`` that runs decode -> resize -> crop -> encode on a
thread pool (PIL releases the GIL in its codecs) and writes records strictly in list order through the indexed writer.  A ``prefix`` that
matches several ``.lst`` files (``prefix_train.lst``, ``prefix_0.lst`` ...) packs each of them.  ``tools/im2rec.cc`` is the native
pass-through packer for the same list format.

  python tools/im2rec.py --list --recursive --train-ratio 0.9 data/train images/     # data/train_train.lst + data/train_val.lst
  python tools/im2rec.py --resize 256 --quality 90 --num-thread 8 data/train images/ # every data/train*.lst -> .rec + .idx
"""
import argparse
import os
import random
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Lister:
    """Enumerates images under ``root`` and writes the tab-separated list files  index<TAB>label[<TAB>label...]<TAB>relative/path."""

    def __init__(self, root, exts, recursive):
        self.root, self.exts, self.recursive = root, tuple(e.lower() for e in exts), recursive

    def items(self):
        """(relative path, label).  recursive: every directory that holds images is a class, numbered in sorted walk order; flat: label 0."""
        out = []
        if self.recursive:
            label_of = {}
            for dirpath, dirs, files in os.walk(self.root, followlinks=True):
                dirs.sort()
                imgs = sorted(f for f in files if f.lower().endswith(self.exts))
                if not imgs:
                    continue
                label = label_of.setdefault(os.path.relpath(dirpath, self.root), len(label_of))
                out += [(os.path.relpath(os.path.join(dirpath, f), self.root), label) for f in imgs]
            for d, l in sorted(label_of.items(), key=lambda kv: kv[1]):
                print(d, l)
        else:
            out = [(f, 0) for f in sorted(os.listdir(self.root)) if os.path.isfile(os.path.join(self.root, f)) and f.lower().endswith(self.exts)]
        return [(i, path, label) for i, (path, label) in enumerate(out)]

    @staticmethod
    def write(path, items):
        with open(path, "w") as f:
            for i, p, label in items:
                f.write("%d\t%f\t%s\n" % (i, label, p))

    def run(self, prefix, shuffle, chunks, train_ratio, test_ratio):
        items = self.items()
        if shuffle:
            random.Random(100).shuffle(items)
        n = len(items)
        per = (n + chunks - 1) // chunks if n else 0
        written = []
        for c in range(chunks):
            chunk = items[c * per:(c + 1) * per]
            tag = "_%d" % c if chunks > 1 else ""
            n_test = int(len(chunk) * test_ratio)
            n_train = int(len(chunk) * train_ratio)
            if train_ratio == 1.0 and test_ratio == 0:
                parts = [("", chunk)]
            else:
                parts = [("_test", chunk[:n_test]), ("_train", chunk[n_test:n_test + n_train]), ("_val", chunk[n_test + n_train:])]
            for suffix, part in parts:
                if part:
                    Lister.write(prefix + tag + suffix + ".lst", part)
                    written.append((prefix + tag + suffix + ".lst", len(part)))
        return written


def read_list(path):
    with open(path) as f:
        for n, line in enumerate(f):
            parts = [p.strip() for p in line.strip().split("\t")]
            if len(parts) < 3:
                if line.strip():
                    print("lst line %d should have at least 3 tab-separated fields, got %r" % (n + 1, line.strip()), file=sys.stderr)
                continue
            try:
                yield int(parts[0]), [float(x) for x in parts[1:-1]], parts[-1]
            except ValueError as e:
                print("lst line %d: %s" % (n + 1, e), file=sys.stderr)


class Packer:
    def __init__(self, args):
        self.a = args

    def encode(self, item):
        """-> (index, packed record bytes) or (index, None) when the image cannot be read."""
        from geomx_b200 import recordio
        idx, labels, rel = item
        a = self.a
        header = recordio.IRHeader(0, labels if (a.pack_label or len(labels) > 1) else labels[0], idx, 0)
        full = os.path.join(a.root, rel)
        try:
            if a.pass_through:
                with open(full, "rb") as f:
                    return idx, recordio.pack(header, f.read())
            import numpy as np
            from PIL import Image
            img = Image.open(full)
            if a.color == 1:
                img = img.convert("RGB")
            elif a.color == 0:
                img = img.convert("L")                      # -1: keep what the file has (alpha included)
            if a.center_crop and img.size[0] != img.size[1]:
                w, h = img.size
                m = min(w, h)
                img = img.crop(((w - m) // 2, (h - m) // 2, (w - m) // 2 + m, (h - m) // 2 + m))
            if a.resize:
                w, h = img.size
                if w > h:
                    img = img.resize((max(1, w * a.resize // h), a.resize), Image.BILINEAR)
                else:
                    img = img.resize((a.resize, max(1, h * a.resize // w)), Image.BILINEAR)
            return idx, recordio.pack_img(header, np.asarray(img), quality=a.quality, img_fmt=a.encoding)
        except Exception as e:                               # a broken image must not stop a million-image job
            print("skipping %s: %s" % (full, e), file=sys.stderr)
            return idx, None

    def run(self, lst):
        from geomx_b200 import recordio
        base = os.path.splitext(lst)[0]
        items = list(read_list(lst))
        rec = recordio.MXIndexedRecordIO(base + ".idx", base + ".rec", "w")
        done = 0
        with ThreadPoolExecutor(max_workers=max(1, self.a.num_thread)) as pool:
            for idx, blob in pool.map(self.encode, items):   # map keeps list order: the .rec is deterministic whatever the thread count
                if blob is not None:
                    rec.write_idx(idx, blob)
                    done += 1
                    if done % 1000 == 0:
                        print("packed", done)
        rec.close()
        print("packed %d of %d images -> %s.rec" % (done, len(items), base))
        return done


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Create an image list or an image-record database.")
    ap.add_argument("prefix", help="prefix of the .lst / .rec / .idx files")
    ap.add_argument("root", help="folder that contains the images")
    g = ap.add_argument_group("list creation")
    g.add_argument("--list", action="store_true", help="create the list file(s) instead of the database")
    g.add_argument("--exts", nargs="+", default=[".jpeg", ".jpg", ".png"])
    g.add_argument("--chunks", type=int, default=1)
    g.add_argument("--train-ratio", type=float, default=1.0)
    g.add_argument("--test-ratio", type=float, default=0.0)
    g.add_argument("--recursive", action="store_true", help="one class per sub-directory")
    g.add_argument("--no-shuffle", dest="shuffle", action="store_false")
    g.add_argument("--shuffle", dest="shuffle", action="store_true")
    g = ap.add_argument_group("database creation")
    g.add_argument("--pass-through", action="store_true", help="store the files as they are")
    g.add_argument("--resize", type=int, default=0, help="resize the shorter edge to this many pixels")
    g.add_argument("--center-crop", action="store_true")
    g.add_argument("--quality", type=int, default=95, help="JPEG quality 1-100 / PNG compression 1-9")
    g.add_argument("--num-thread", type=int, default=1)
    g.add_argument("--color", type=int, default=1, choices=[-1, 0, 1])
    g.add_argument("--encoding", default=".jpg", choices=[".jpg", ".png"])
    g.add_argument("--pack-label", action="store_true", help="store multi-dimensional labels (detection lists)")
    ap.set_defaults(shuffle=True)
    a = ap.parse_args(argv)
    a.prefix, a.root = os.path.abspath(a.prefix), os.path.abspath(a.root)
    return a


def main(argv=None):
    a = parse(argv)
    if a.list:
        for path, n in Lister(a.root, a.exts, a.recursive).run(a.prefix, a.shuffle, max(1, a.chunks), a.train_ratio, a.test_ratio):
            print("wrote %d entries -> %s" % (n, path))
        return 0
    folder, stem = os.path.dirname(a.prefix) or ".", os.path.basename(a.prefix)
    lists = sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.startswith(stem) and f.endswith(".lst"))
    if not lists:
        print("no list file matches %s*.lst — run with --list first" % a.prefix, file=sys.stderr)
        return 1
    packer = Packer(a)
    for lst in lists:
        packer.run(lst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
