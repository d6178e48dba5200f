#!/usr/bin/env python
"""Hierarchical Frequency Aggregation: K1 local Adam steps, then a synchronisation round in which every worker pushes ``params / num_local_workers``;
every K2-th round the parties' progress is merged globally (MXNET_KVSTORE_USE_HFA=1, MXNET_KVSTORE_HFA_K1, MXNET_KVSTORE_HFA_K2)."""
import os

from common import checkpointing, Progress, accuracy, build_net, make_loaders, make_parser, mx, pick_context, worker_slice


def main():
    args = make_parser().parse_args()
    use_hfa, k1, k2 = (int(os.getenv(k, 0)) for k in ("MXNET_KVSTORE_USE_HFA", "MXNET_KVSTORE_HFA_K1", "MXNET_KVSTORE_HFA_K2"))
    assert use_hfa == 1 and k1 >= 1 and k2 >= 1, "MXNET_KVSTORE_USE_HFA / _K1 / _K2 are not properly set"
    ts_on = int(os.getenv("ENABLE_INTER_TS", 0)) or int(os.getenv("ENABLE_INTRA_TS", 0))
    ctx = pick_context(args.cpu)
    net = build_net(ctx, args.batch_size)
    bind_kv = checkpointing(net, args)
    kv = mx.kv.create("dist_sync")
    bind_kv(kv)
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
            trainer.step(X.shape[0])
            mx.nd.waitall()
            prog.it += 1
            if prog.it % k1 == 0:
                for idx, p in enumerate(params):
                    kv.push(idx, p.data() / kv.num_workers, priority=-idx)
                    kv.pull(idx, p.data(), priority=-idx)
                    if ts_on:
                        mx.nd.waitall()
                mx.nd.waitall()
            if prog.it % (k1 * k2) == 0:
                prog.log(epoch, accuracy(test, net, ctx))
            if args.max_iters and prog.it >= args.max_iters:
                return


if __name__ == "__main__":
    main()
