"""HiPS job description → per-process environments → local / ssh / mpi launch.

    python -m geomx_b200.tracker.launch -n 2 -s 1 python examples/cnn.py --cpu                       # single tier: scheduler + 1 server + 2 workers
    python -m geomx_b200.tracker.launch --parties 2 -n 2 --global-servers 1 python examples/cnn.py   # two tiers: the reference's 12 processes
    python -m geomx_b200.tracker.launch --launcher ssh -H hosts --parties 2 -n 2 python examples/cnn.py

Roles and variables follow ``3rdparty/ps-lite/src/postoffice.cc:18-58`` (``DMLC_ROLE``, ``DMLC_ROLE_GLOBAL``, ``DMLC_NUM_{WORKER,SERVER,
GLOBAL_WORKER,GLOBAL_SERVER,ALL_WORKER}``, ``DMLC_ROLE_MASTER_WORKER``, ``DMLC_ENABLE_CENTRAL_WORKER``, ``DMLC_PS_ROOT_{URI,PORT}``,
``DMLC_PS_GLOBAL_ROOT_{URI,PORT}``).  With ``--launcher ssh`` parties are placed on consecutive hosts of the host file (central party on
the first one), which is the multi-datacentre layout HiPS exists for; ``mpi`` prints/executes one ``mpirun`` per process group.

Cluster schedulers (capability parity with the reference's dmlc-core tracker back-ends ``mpi / sge / slurm / kubernetes``,
3rdparty/dmlc-core/tracker/dmlc_tracker/{mpi,sge,slurm,kubernetes}.py; yarn / mesos are Hadoop-era and not rebuilt): every back-end is a
function ``(Proc, argv, env, cwd) -> command`` over the same process table, so a HiPS job has the identical layout whatever starts it:
``slurm`` = one ``srun`` step per process pinned to the process's host, ``sge`` = one ``qsub`` of a generated job script per process,
``kubernetes`` = ``--dry-run`` prints (and a real run ``kubectl apply``s) one manifest with a headless Service + Pod per process, the
rendezvous addresses rewritten to the Services' DNS names.
"""
from __future__ import annotations

import argparse
import os
import shlex
import subprocess
import sys
from dataclasses import dataclass, field

BOOT = "import geomx_b200"     # non-worker roles: importing the package runs the server / scheduler loop (kvstore_server.py)


@dataclass
class Proc:
    name: str
    env: dict
    is_worker: bool
    host: str = "127.0.0.1"


@dataclass
class HipsJob:
    workers_per_party: int = 2
    servers: int = 1                 # single-tier only
    parties: int = 0                 # 0 = single tier
    global_servers: int = 1
    base_port: int = 9092
    hosts: list = field(default_factory=lambda: ["127.0.0.1"])
    extra_env: dict = field(default_factory=dict)
    central_worker: bool = False

    def _host(self, i):
        return self.hosts[i % len(self.hosts)]

    def processes(self):
        procs = []
        if self.parties <= 0:
            root = self._host(0)
            base = {"DMLC_PS_ROOT_URI": root, "DMLC_PS_ROOT_PORT": self.base_port, "DMLC_NUM_SERVER": self.servers,
                    "DMLC_NUM_WORKER": self.workers_per_party, "DMLC_NUM_ALL_WORKER": self.workers_per_party}
            procs.append(Proc("scheduler", dict(base, DMLC_ROLE="scheduler"), False, root))
            for s in range(self.servers):
                procs.append(Proc("server%d" % s, dict(base, DMLC_ROLE="server"), False, self._host(s)))
            for w in range(self.workers_per_party):
                procs.append(Proc("worker%d" % w, dict(base, DMLC_ROLE="worker", GEOMX_WORKER_INDEX=w), True, self._host(w)))
            return procs
        allw = self.parties * self.workers_per_party
        central = self._host(0)
        genv = {"DMLC_PS_GLOBAL_ROOT_URI": central, "DMLC_PS_GLOBAL_ROOT_PORT": self.base_port, "DMLC_NUM_GLOBAL_SERVER": self.global_servers,
                "DMLC_NUM_GLOBAL_WORKER": self.parties}
        cenv = {"DMLC_PS_ROOT_URI": central, "DMLC_PS_ROOT_PORT": self.base_port + 1, "DMLC_NUM_SERVER": self.global_servers, "DMLC_NUM_WORKER": 1,
                "DMLC_NUM_ALL_WORKER": allw}
        procs.append(Proc("global_scheduler", dict(genv, DMLC_ROLE_GLOBAL="global_scheduler"), False, central))
        for g in range(self.global_servers):
            procs.append(Proc("global_server%d" % g, dict(genv, **cenv, DMLC_ROLE_GLOBAL="global_server", DMLC_ROLE="server",
                                                           DMLC_ENABLE_CENTRAL_WORKER=int(self.central_worker)), False, central))
        procs.append(Proc("central_scheduler", dict(cenv, DMLC_ROLE="scheduler"), False, central))
        procs.append(Proc("master_worker", dict(cenv, DMLC_ROLE="worker", DMLC_ROLE_MASTER_WORKER=1), True, central))
        idx = 0
        for p in range(self.parties):
            host = self._host(1 + p) if len(self.hosts) > 1 else central
            penv = {"DMLC_PS_ROOT_URI": host, "DMLC_PS_ROOT_PORT": self.base_port + 2 + p, "DMLC_NUM_SERVER": 1,
                    "DMLC_NUM_WORKER": self.workers_per_party, "DMLC_NUM_ALL_WORKER": allw}
            procs.append(Proc("party%d_scheduler" % p, dict(penv, DMLC_ROLE="scheduler"), False, host))
            procs.append(Proc("party%d_server" % p, dict(genv, **penv, DMLC_ROLE="server"), False, host))
            for w in range(self.workers_per_party):
                procs.append(Proc("party%d_worker%d" % (p, w), dict(penv, DMLC_ROLE="worker", GEOMX_WORKER_INDEX=idx), True, host))
                idx += 1
        return procs

    def command_lines(self, worker_cmd, python=sys.executable):
        """[(Proc, argv, env)] — what each launcher executes."""
        out = []
        for p in self.processes():
            env = {k: str(v) for k, v in {**self.extra_env, **p.env}.items()}
            argv = list(worker_cmd) if p.is_worker else [python, "-c", BOOT]
            out.append((p, argv, env))
        return out


def _ssh_line(host, argv, env, cwd):
    exports = " ".join("%s=%s" % (k, shlex.quote(v)) for k, v in env.items())
    return ["ssh", "-o", "StrictHostKeyChecking=no", host, "cd %s && env %s %s" % (shlex.quote(cwd), exports, " ".join(shlex.quote(a) for a in argv))]


def _mpi_line(host, argv, env):
    xs = []
    for k, v in env.items():
        xs += ["-x", "%s=%s" % (k, v)]
    return ["mpirun", "-n", "1", "--host", host] + xs + list(argv)


def _slurm_line(p, argv, env, cwd):
    exports = ",".join(["ALL"] + ["%s=%s" % kv for kv in env.items()])
    host = [] if p.host in ("127.0.0.1", "localhost") else ["--nodelist", p.host]
    return ["srun", "--nodes=1", "--ntasks=1", "--job-name", "hips-" + p.name, "--chdir", cwd, "--export", exports] + host + list(argv)


def _sge_script(p, argv, env, cwd):
    """Job script of one role for Sun Grid Engine (submitted with ``qsub -sync y`` so that the launcher sees the exit code)."""
    lines = ["#!/bin/bash", "#$ -S /bin/bash", "#$ -N hips-" + p.name, "#$ -cwd", "#$ -j y"]
    if p.host not in ("127.0.0.1", "localhost"):
        lines.append("#$ -l hostname=" + p.host)
    lines += ["export %s=%s" % (k, shlex.quote(v)) for k, v in env.items()]
    lines += ["cd " + shlex.quote(cwd), "exec " + " ".join(shlex.quote(a) for a in argv), ""]
    return "\n".join(lines)


def kubernetes_manifest(job: HipsJob, worker_cmd, image="geomx-b200:latest", namespace="default", python="python"):
    """One YAML document per process: a headless Service (stable DNS name for the rendezvous) and a Pod.  Scheduler / root addresses are
    rewritten from host names to Service names, so nothing in the job depends on Pod IPs."""
    lines_ = job.command_lines(worker_cmd, python=python)
    svc = {}                                              # rendezvous host -> Service name of the process that listens there
    for p, argv, env in lines_:
        name = "hips-" + p.name.replace("_", "-")
        if p.name == "global_scheduler":
            svc[("g", p.host, env["DMLC_PS_GLOBAL_ROOT_PORT"])] = name
        elif p.name.endswith("scheduler"):
            svc[("l", p.host, env["DMLC_PS_ROOT_PORT"])] = name
    docs = []
    for p, argv, env in lines_:
        name = "hips-" + p.name.replace("_", "-")
        e = dict(env)
        gk = ("g", e.get("DMLC_PS_GLOBAL_ROOT_URI"), e.get("DMLC_PS_GLOBAL_ROOT_PORT"))
        lk = ("l", e.get("DMLC_PS_ROOT_URI"), e.get("DMLC_PS_ROOT_PORT"))
        if gk in svc:
            e["DMLC_PS_GLOBAL_ROOT_URI"] = svc[gk]
        if lk in svc:
            e["DMLC_PS_ROOT_URI"] = svc[lk]
        e["DMLC_NODE_HOST"] = name
        envs = "\n".join("        - {name: %s, value: %s}" % (k, _yq(v)) for k, v in sorted(e.items()))
        cmd = ", ".join(_yq(a) for a in argv)
        docs.append(("apiVersion: v1\nkind: Service\nmetadata: {name: %s, namespace: %s}\nspec:\n  clusterIP: None\n  selector: {hips-role: %s}\n"
                     "  ports: [{name: ps, port: 9000}]\n---\napiVersion: v1\nkind: Pod\nmetadata:\n  name: %s\n  namespace: %s\n  labels: {hips-role: %s, "
                     "hips-job: hips}\nspec:\n  restartPolicy: Never\n  hostname: %s\n  containers:\n    - name: main\n      image: %s\n"
                     "      command: [%s]\n      env:\n%s\n") % (name, namespace, name, name, namespace, name, name, image, cmd, envs)
                    + ("      resources: {limits: {nvidia.com/gpu: 1}}\n" if p.is_worker else ""))
    return "---\n".join(docs)


def _yq(v):
    return '"%s"' % str(v).replace("\\", "\\\\").replace('"', '\\"')


def launch(job: HipsJob, worker_cmd, launcher="local", log_dir=None, dry_run=False, timeout=None):
    """Start every process of ``job``; returns the worst exit code (``dry_run``: the command lines instead)."""
    lines = job.command_lines(worker_cmd)
    cwd = os.getcwd()
    plan = []
    for p, argv, env in lines:
        if launcher == "ssh" and p.host not in ("127.0.0.1", "localhost"):
            plan.append((p, _ssh_line(p.host, argv, env, cwd), None))
        elif launcher == "mpi":
            plan.append((p, _mpi_line(p.host, argv, env), None))
        elif launcher == "slurm":
            plan.append((p, _slurm_line(p, argv, env, cwd), None))
        elif launcher == "sge":
            script_dir = log_dir or os.path.join(cwd, ".hips_sge")
            path = os.path.join(script_dir, p.name + ".sh")
            if not dry_run:
                os.makedirs(script_dir, exist_ok=True)
                with open(path, "w") as f:
                    f.write(_sge_script(p, argv, env, cwd))
            plan.append((p, ["qsub", "-sync", "y", path], None))
        else:
            plan.append((p, argv, env))
    if launcher == "kubernetes":
        manifest = kubernetes_manifest(job, worker_cmd)
        if dry_run:
            return [("manifest", ["kubectl", "apply", "-f", "-"], {"MANIFEST": manifest})]
        return subprocess.run(["kubectl", "apply", "-f", "-"], input=manifest.encode()).returncode
    if dry_run:
        return [(p.name, cmd, env) for p, cmd, env in plan]
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
    procs = []
    for p, cmd, env in plan:
        e = dict(os.environ)
        e.pop("RANK", None); e.pop("WORLD_SIZE", None)
        if env:
            e.update(env)
        out = open(os.path.join(log_dir, p.name + ".log"), "w") if log_dir else None
        procs.append((p, subprocess.Popen(cmd, env=e, stdout=out, stderr=subprocess.STDOUT if out else None), out))
    # fail fast: a role that dies at start-up would leave every other process waiting at the rendezvous forever
    import time
    rc, t0 = 0, time.time()
    alive = list(procs)
    while alive:
        nxt = []
        for item in alive:
            r = item[1].poll()
            if r is None:
                nxt.append(item)
            elif r != 0 and rc == 0:
                rc = r
                print("process %s exited with code %d - stopping the job" % (item[0].name, r), file=sys.stderr)
        alive = nxt
        if alive and timeout is not None and time.time() - t0 > timeout:
            rc = rc or 124
        if rc != 0:
            for _, q, _ in alive:
                q.kill()
            break
        if alive:
            time.sleep(0.2)
    for _, q, out in procs:
        q.wait()
        if out:
            out.close()
    return rc


def main(argv=None):
    ap = argparse.ArgumentParser(description="launch a (hierarchical) parameter-server job")
    ap.add_argument("-n", "--num-workers", type=int, default=2, help="workers (per party when --parties > 0)")
    ap.add_argument("-s", "--num-servers", type=int, default=1, help="servers of a single-tier job")
    ap.add_argument("--parties", type=int, default=0, help="number of participating parties (0: single tier)")
    ap.add_argument("--global-servers", type=int, default=1)
    ap.add_argument("--central-worker", action="store_true")
    ap.add_argument("--launcher", default="local", choices=["local", "ssh", "mpi", "slurm", "sge", "kubernetes"])
    ap.add_argument("-H", "--hostfile")
    ap.add_argument("--base-port", type=int, default=9092)
    ap.add_argument("--log-dir")
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE for every process (repeatable), e.g. ENABLE_P3=1")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("command", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    hosts = [l.strip() for l in open(a.hostfile) if l.strip() and not l.startswith("#")] if a.hostfile else ["127.0.0.1"]
    job = HipsJob(a.num_workers, a.num_servers, a.parties, a.global_servers, a.base_port, hosts, dict(kv.split("=", 1) for kv in a.env),
                  a.central_worker)
    cmd = a.command[1:] if a.command and a.command[0] == "--" else a.command
    if not cmd:
        ap.error("missing worker command")
    res = launch(job, cmd, a.launcher, a.log_dir, a.dry_run)
    if a.dry_run and a.launcher == "kubernetes":
        print(res[0][2]["MANIFEST"])
        return 0
    if a.dry_run:
        for name, c, env in res:
            print(name, ":", " ".join(shlex.quote(x) for x in c), "| env:", " ".join("%s=%s" % kv for kv in sorted((env or {}).items())))
        return 0
    return res


if __name__ == "__main__":
    sys.exit(main())
