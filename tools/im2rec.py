#!/usr/bin/env python
"""Create ``.lst`` / ``.rec`` (+ ``.idx``) image-record files from an image folder (parity: ``tools/im2rec.py`` of the reference).

  python tools/im2rec.py --list prefix root        # walk root/<class>/*.{jpg,png} -> prefix.lst
  python tools/im2rec.py prefix root               # prefix.lst -> prefix.rec + prefix.idx  (optionally --resize N --quality Q)
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

EXTS = (".jpg", ".jpeg", ".png")


def make_list(args):
    classes = sorted(d for d in os.listdir(args.root) if os.path.isdir(os.path.join(args.root, d)))
    items = []
    for label, c in enumerate(classes):
        for dirpath, _, files in os.walk(os.path.join(args.root, c)):
            for f in sorted(files):
                if f.lower().endswith(EXTS):
                    items.append((os.path.relpath(os.path.join(dirpath, f), args.root), label))
    if args.shuffle:
        random.Random(100).shuffle(items)
    with open(args.prefix + ".lst", "w") as out:
        for i, (path, label) in enumerate(items):
            out.write("%d\t%f\t%s\n" % (i, label, path))
    print("wrote %d entries, %d classes -> %s.lst" % (len(items), len(classes), args.prefix))


def make_rec(args):
    import numpy as np
    from PIL import Image
    from geomx_b200 import recordio
    rec = recordio.MXIndexedRecordIO(args.prefix + ".idx", args.prefix + ".rec", "w")
    n = 0
    for line in open(args.prefix + ".lst"):
        parts = line.strip().split("\t")
        if len(parts) < 3:
            continue
        idx, labels, path = int(parts[0]), [float(x) for x in parts[1:-1]], parts[-1]
        full = os.path.join(args.root, path)
        header = recordio.IRHeader(0, labels[0] if len(labels) == 1 else labels, idx, 0)
        if args.pass_through:
            rec.write_idx(idx, recordio.pack(header, open(full, "rb").read()))
        else:
            img = Image.open(full).convert("RGB" if args.color else "L")
            if args.resize:
                w, h = img.size
                s = args.resize / float(min(w, h))
                img = img.resize((max(1, int(w * s)), max(1, int(h * s))))
            rec.write_idx(idx, recordio.pack_img(header, np.asarray(img), quality=args.quality, img_fmt=args.encoding))
        n += 1
    rec.close()
    print("packed %d records -> %s.rec" % (n, args.prefix))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix"); ap.add_argument("root")
    ap.add_argument("--list", action="store_true"); ap.add_argument("--shuffle", action="store_true")
    ap.add_argument("--resize", type=int, default=0); ap.add_argument("--quality", type=int, default=95)
    ap.add_argument("--encoding", default=".jpg"); ap.add_argument("--color", type=int, default=1)
    ap.add_argument("--pass-through", action="store_true")
    a = ap.parse_args()
    (make_list if a.list else make_rec)(a)
