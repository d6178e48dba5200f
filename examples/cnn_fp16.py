#!/usr/bin/env python
"""FP16 transport: every tensor crosses the parameter-server tiers as float16 (half the traffic); the update is a local fp32 Adam."""
import time

from common import checkpointing, Progress, accuracy, build_net, make_loaders, make_parser, mx, pick_context, worker_slice


def run(args, low_precision_for):
    ctx = pick_context(args.cpu)
    net = build_net(ctx, args.batch_size)
    bind_kv = checkpointing(net, args)
    kv = mx.kv.create("dist_sync")
    bind_kv(kv)
    if getattr(args, "bisparse_compression_ratio", None) and (kv.is_master_worker or getattr(kv, "configures_servers", False)):
        kv.set_gradient_compression({"type": "bsc", "threshold": args.bisparse_compression_ratio})
    time.sleep(1)
    trainer = mx.gluon.Trainer(net.collect_params(), optimizer=mx.optimizer.Adam(learning_rate=args.learning_rate), kvstore=None, update_on_kvstore=False)
    loss_fn = mx.gluon.loss.SoftmaxCrossEntropyLoss()
    params = list(net.collect_params().values())
    for idx, p in enumerate(params):
        buf = p.data().astype("float16") if low_precision_for(p.data()) else p.data().copy()
        kv.init(idx, buf)
        if kv.is_master_worker:
            continue
        kv.pull(idx, buf)
        p.set_data(buf.astype("float32"))
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
                g = p.grad().astype("float16") if low_precision_for(p.grad()) else p.grad().copy()
                kv.push(idx, g / n, priority=-idx)
                kv.pull(idx, g, priority=-idx)
                p.grad()[:] = g.astype("float32")
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
    run(make_parser().parse_args(), lambda arr: True)
