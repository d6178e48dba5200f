#!/usr/bin/env python
"""Vanilla HiPS training of the demo CNN: FSA (dist_sync), MixedSync (-ms, dist_async) or MixedSync + DCASGD (-dc).
The optimizer runs on the global parameter server; workers push ``grad / batch`` and pull fresh weights, key by key with priority -idx."""
import os
import time

from common import checkpointing, Progress, accuracy, build_net, configures_servers, make_loaders, make_parser, mx, pick_context, worker_slice


def main():
    args = make_parser(extra=("mixed",)).parse_args()
    ctx = pick_context(args.cpu)
    ts_on = int(os.getenv("ENABLE_INTER_TS", 0)) or int(os.getenv("ENABLE_INTRA_TS", 0))
    net = build_net(ctx, args.batch_size)
    bind_kv = checkpointing(net, args)

    kv = mx.kv.create("dist_async" if (args.mixed_sync or args.dcasgd) else "dist_sync")

    bind_kv(kv)
    if configures_servers(kv):
        opt = mx.optimizer.DCASGD(learning_rate=args.learning_rate) if args.dcasgd else mx.optimizer.Adam(learning_rate=args.learning_rate)
        kv.set_optimizer(opt)
    time.sleep(1)                                   # let the configuration reach every server
    loss_fn = mx.gluon.loss.SoftmaxCrossEntropyLoss()
    params = list(net.collect_params().values())
    for idx, p in enumerate(params):
        kv.init(idx, p.data())
        if not kv.is_master_worker:
            kv.pull(idx, p.data())
    mx.nd.waitall()
    if kv.is_master_worker:
        return

    train, test = make_loaders(args.batch_size, kv.num_all_workers, worker_slice(args, kv), args.data_dir, args.split_by_class)
    prog = Progress()
    print("Start training on %d workers, my rank is %d." % (kv.num_all_workers, kv.rank), flush=True)
    for epoch in range(args.epoch):
        for X, y in train:
            X, y = X.as_in_context(ctx), y.as_in_context(ctx)
            with mx.autograd.record():
                l = loss_fn(net(X), y)
            l.backward()
            n = X.shape[0]
            for idx, p in enumerate(params):
                if p.grad_req == "null":
                    continue
                kv.push(idx, p.grad() / n, priority=-idx)
                kv.pull(idx, p.data(), priority=-idx)
                if ts_on:
                    mx.nd.waitall()
            mx.nd.waitall()
            prog.it += 1
            if args.eval_every and prog.it % args.eval_every == 0:
                prog.log(epoch, accuracy(test, net, ctx))
            if args.max_iters and prog.it >= args.max_iters:
                return


if __name__ == "__main__":
    main()
