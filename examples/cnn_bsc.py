#!/usr/bin/env python
"""Bi-Sparse compression: the master worker enables ``bsc`` on the servers; workers push gradients, pull the (sparsified) global aggregate
and apply Adam locally (``Trainer.step(num_all_workers)``)."""
import time

from common import checkpointing, Progress, accuracy, build_net, configures_servers, make_loaders, make_parser, mx, pick_context, worker_slice


def main():
    args = make_parser(extra=("bcr",)).parse_args()
    assert 0 < args.bisparse_compression_ratio < 1, "bisparse_compression_ratio is not properly set"
    ctx = pick_context(args.cpu)
    net = build_net(ctx, args.batch_size)
    bind_kv = checkpointing(net, args)
    kv = mx.kv.create("dist_sync")
    bind_kv(kv)
    if configures_servers(kv):
        kv.set_gradient_compression({"type": "bsc", "threshold": args.bisparse_compression_ratio})
    time.sleep(1)
    trainer = mx.gluon.Trainer(net.collect_params(), optimizer=mx.optimizer.Adam(learning_rate=args.learning_rate), kvstore=None, update_on_kvstore=False)
    loss_fn = mx.gluon.loss.SoftmaxCrossEntropyLoss()
    params = list(net.collect_params().values())
    for idx, p in enumerate(params):
        kv.init(idx, p.data().copy())
        if not kv.is_master_worker:
            kv.pull(idx, p.data())
    mx.nd.waitall()
    if kv.is_master_worker:
        return
    train, test = make_loaders(args.batch_size, kv.num_all_workers, worker_slice(args, kv), args.data_dir, args.split_by_class)
    prog = Progress()
    for epoch in range(args.epoch):
        for X, y in train:
            X, y = X.as_in_context(ctx), y.as_in_context(ctx)
            with mx.autograd.record():
                l = loss_fn(net(X), y)
            l.backward()
            n = X.shape[0]
            for idx, p in enumerate(params):
                kv.push(idx, p.grad() / n, priority=-idx)
                kv.pull(idx, p.grad(), priority=-idx)        # aggregated gradients come back
            mx.nd.waitall()
            trainer.step(kv.num_all_workers)
            for p in params:
                p.zero_grad()
            prog.it += 1
            if args.eval_every and prog.it % args.eval_every == 0:
                prog.log(epoch, accuracy(test, net, ctx))
            if args.max_iters and prog.it >= args.max_iters:
                return


if __name__ == "__main__":
    main()
