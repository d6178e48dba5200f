"""Server / scheduler bootstrap executed on ``import geomx_b200``.

Parity: ``python/mxnet/kvstore_server.py:30-89``: if the process role is not *worker* (or it is the global scheduler)
create a ``dist`` kvstore, run the server loop, then ``sys.exit()``.  The controller unpickles an optimizer shipped
by ``KVStore.set_optimizer`` (``kController``) and installs it as the server's updater.  Here the native server
(``csrc/hips/server.cc``) runs natively-implemented optimizers without Python; arbitrary pickled optimizers are
executed through a host callback on this (main) thread — the equivalent of the reference's ``Executor`` hand-off
(``kvstore_dist_server.h:109-168``)."""
from __future__ import annotations

import logging
import os
import pickle
import sys

__all__ = ["KVStoreServer", "_init_kvstore_server_module"]


def _role():
    return os.environ.get("DMLC_ROLE", ""), os.environ.get("DMLC_ROLE_GLOBAL", "")


def is_worker_node():
    return _role()[0] in ("", "worker")


def is_global_scheduler_node():
    return _role()[1] == "global_scheduler"


class KVStoreServer:
    def __init__(self, kvstore):
        self.kvstore = kvstore
        self.handle = kvstore
        self.init_logging = False

    def _controller(self):
        def server_controller(cmd_id, cmd_body):
            if not self.init_logging:
                head = "%(asctime)-15s Server[" + str(self.kvstore.rank) + "] %(message)s"
                logging.basicConfig(level=logging.DEBUG, format=head)
                self.init_logging = True
            if cmd_id == 0:
                body = cmd_body if isinstance(cmd_body, bytes) else cmd_body.encode("latin1")
                optimizer = pickle.loads(body)
                self.kvstore.set_optimizer(optimizer)
            else:
                print("server %d, unknown command (%d, %s)" % (self.kvstore.rank, cmd_id, cmd_body))
        return server_controller

    def run(self):
        self.kvstore.run_server(self._controller())


def _init_kvstore_server_module():
    role, grole = _role()
    if os.environ.get("GEOMX_NO_SERVER_BOOTSTRAP", "0") == "1":
        return
    if (role and role != "worker") or grole == "global_scheduler":
        from . import kvstore as kvs
        kv = kvs.create("dist")
        server = KVStoreServer(kv)
        server.run()
        sys.exit()
