"""``mx.gluon.model_zoo.model_store`` — locating pretrained parameter files (reference: ``python/mxnet/gluon/model_zoo/model_store.py``).

The reference downloads ``<name>-<short hash>.params`` from its S3 repository into ``~/.mxnet/models``.  There is no network here: the store
only RESOLVES files that are already on disk — in ``root`` (default ``$MXNET_HOME/models`` or ``~/.mxnet/models``) under either the
reference's hashed name or plain ``<name>.params`` — and says exactly where it looked otherwise.  ``.params`` files written by MXNet load
unchanged (``docs/migration.md``), so a directory copied from a machine with network access works as a cache."""
import glob
import os

from ...base import MXNetError

__all__ = ["get_model_file", "purge", "short_hash", "check_sha1"]

def _default_root():
    return os.path.join(os.environ.get("MXNET_HOME", os.path.join(os.path.expanduser("~"), ".mxnet")), "models")


def short_hash(name, root=None):
    """The 8-hex-digit tag in the file name of a cached model (``<name>-<tag>.params``, the reference's naming: first digits of the file's
    sha1); raises if no such file is cached."""
    path = get_model_file(name, root)
    base = os.path.basename(path)[len(name):]
    if not (base.startswith("-") and base.endswith(".params")):
        raise ValueError("cached file of %s carries no hash tag: %s" % (name, path))
    return base[1:-len(".params")]


def check_sha1(filename, sha1_hash):
    """True if the sha1 of ``filename`` starts with / equals ``sha1_hash`` (``mxnet.gluon.utils.check_sha1`` semantics)."""
    import hashlib
    h = hashlib.sha1()
    with open(filename, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest().startswith(sha1_hash.lower())


def get_model_file(name, root=None):
    """Path of the parameter file of ``name`` inside ``root`` — ``<name>-<tag>.params`` (a tagged file is verified against its tag) or
    ``<name>.params``; raises with the places searched if there is none."""
    root = os.path.expanduser(root or _default_root())
    tagged = sorted(glob.glob(os.path.join(root, name + "-*.params")))
    for c in tagged:
        tag = os.path.basename(c)[len(name) + 1:-len(".params")]
        if len(tag) == 8 and all(ch in "0123456789abcdef" for ch in tag.lower()) and not check_sha1(c, tag):
            raise MXNetError("%s does not match the hash in its name (corrupt or truncated copy)" % c)
        return c
    plain = os.path.join(root, name + ".params")
    if os.path.isfile(plain):
        return plain
    raise MXNetError("no parameter file for model %r under %s (looked for %s-<hash>.params and %s.params).  Pretrained weights cannot be "
                     "downloaded here — copy the .params file (MXNet's format is read as is) into that directory, or pass root=..."
                     % (name, root, name, name))


def purge(root=None):
    """Remove every cached ``.params`` file under ``root``."""
    root = os.path.expanduser(root or _default_root())
    for f in glob.glob(os.path.join(root, "*.params")):
        os.remove(f)
