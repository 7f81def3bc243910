#!/usr/bin/env python
"""Lattice-free MMI (chain) training of the BLSTM acoustic model on B200 (reference bin/train_chain.py).

Same flags as the reference script (bin/train_chain.py:60-83).  The loop body is :244-308: 3x frame
subsampling with the epoch-dependent shift, model forward, the chain objective per utterance
(``ops.ChainObjtiveFunction``), Noam learning rate, clip, Adam(amsgrad), rank-0 checkpoints
``chain.model.<i>.tar``.  ``-synthetic N`` supplies N seeded utterances, a synthetic denominator FST
(``-den_states``) and synthetic time-constrained numerator FSTs.  Without it the corpus of ``-data`` (zip of wavs +
pdf-id label file) is read, ``-den_fst`` takes Kaldi's den.fst (OpenFst binary or fstprint text) and the numerator
graph of an utterance is built from its pdf alignment with a boundary tolerance (``-tolerance``); the reference's
phone-level proto-supervision (0.trans_mdl, tree; :167-202,262-272) needs Kaldi's tree / topology.
``-per_utt_loss 1`` keeps the reference's one-call-per-utterance loop; the default batches the B
calls into one C-ABI call (identical numbers, tests/test_gpu_fb.py).
"""
import argparse
import os
import time
import zlib

import numpy as np
import torch as th

import _common
from _common import pkdist
from pykaldi2_b200 import graphs, pipeline, synth
from pykaldi2_b200.data.dataloader import SyntheticWaveDataset, WaveDataloader
from pykaldi2_b200.data.speech_dataset import SpeechDataset
from pykaldi2_b200 import chain_supervision
from pykaldi2_b200.data import fbank as fb
from pykaldi2_b200.models import lstm
from pykaldi2_b200.reader import kaldi_io
from pykaldi2_b200.ops import ops
from pykaldi2_b200.utils import utils


def SupervisionOptions():
    """kaldi_chain.SupervisionOptions as the reference sets it (bin/train_chain.py:184-188)."""
    return chain_supervision.SupervisionOptions(left_tolerance=5, right_tolerance=5, frame_subsampling_factor=3,
                                                convert_to_pdfs=True)


def _first_existing(*paths):
    for p in paths:
        if p and os.path.isfile(p):
            return p
    return None


def load_kaldi_assets(args):
    """The reference's Kaldi inputs (bin/train_chain.py:162-181), binary as Kaldi writes them or in text form: the
    alignment model's transition model (<ali_dir>/final.mdl[.txt]), the chain transition model
    (<chain_dir>/0.trans_mdl[.txt]) and tree (<chain_dir>/tree[.txt]).  Returns None when -ali_dir / -chain_dir are
    not given."""
    if not (args.ali_dir and args.chain_dir):
        return None
    ali = _first_existing(args.ali_dir + "/final.mdl.txt", args.ali_dir + "/final.mdl")
    ctm = _first_existing(args.chain_dir + "/0.trans_mdl.txt", args.chain_dir + "/0.trans_mdl")
    tree = _first_existing(args.chain_dir + "/tree.txt", args.chain_dir + "/tree")
    if not (ali and ctm and tree):
        raise SystemExit("train_chain.py: -ali_dir / -chain_dir must hold final.mdl, 0.trans_mdl and tree")
    return {"ali_tm": kaldi_io.read_transition_model(ali), "chain_tm": kaldi_io.read_transition_model(ctm),
            "tree": kaldi_io.read_tree(tree)}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("-config")
    parser.add_argument("-data", help="data yaml file")
    parser.add_argument("-dataPath", default='', type=str, help="path of data files")
    parser.add_argument("-seed_model", default='', help="the seed nerual network model")
    parser.add_argument("-exp_dir", help="the directory to save the outputs")
    parser.add_argument("-transform", help="feature transformation matrix or mvn statistics")
    parser.add_argument("-ali_dir", help="the directory to load trans_model and tree used for alignments")
    parser.add_argument("-lang_dir", help="the lexicon directory to load L.fst")
    parser.add_argument("-chain_dir", help="the directory to load trans_model, tree and den.fst for chain model")
    parser.add_argument("-lr", type=float, default=1e-3, help="set the base learning rate")
    parser.add_argument("-warmup_steps", default=4000, type=int, help="the number of warmup steps to adjust the learning rate")
    parser.add_argument("-xent_regularize", default=0, type=float, help="cross-entropy regularization weight")
    parser.add_argument("-momentum", default=0, type=float, help="set the momentum")
    parser.add_argument("-weight_decay", default=1e-4, type=float, help="set the L2 regularization weight")
    parser.add_argument("-batch_size", default=32, type=int, help="Override the batch size in the config")
    parser.add_argument("-data_loader_threads", default=0, type=int, help="number of workers for data loading")
    parser.add_argument("-max_grad_norm", default=5, type=float, help="max_grad_norm for gradient clipping")
    parser.add_argument("-sweep_size", default=100, type=float, help="process n hours of data per sweep (default:100)")
    parser.add_argument("-num_epochs", default=1, type=int, help="number of training epochs (default:1)")
    parser.add_argument("-anneal_lr_epoch", default=2, type=int, help="start to anneal the learning rate from this epoch")
    parser.add_argument("-anneal_lr_ratio", default=0.5, type=float, help="the ratio to anneal the learning rate ratio")
    parser.add_argument('-print_freq', default=10, type=int, metavar='N', help='print frequency (default: 10)')
    parser.add_argument('-save_freq', default=1000, type=int, metavar='N', help='save model frequency (default: 1000)')
    parser.add_argument('-synthetic', default=0, type=int, help="train on this many seeded synthetic utterances")
    parser.add_argument('-tolerance', default=2, type=int, help="boundary tolerance (output frames) of the alignment-derived numerator graphs")
    parser.add_argument('-den_fst', default='', type=str, help="denominator FST file (OpenFst binary vector/standard or fstprint text); default: synthetic")
    parser.add_argument('-den_states', default=8192, type=int, help="states of the synthetic denominator FST")
    parser.add_argument('-per_utt_loss', default=0, type=int, help="1 = one chain-objective call per utterance as the reference does")
    parser.add_argument('-seed', default=1234, type=int, help="random seed (model init, sampling)")
    parser.add_argument('-max_steps', default=0, type=int)
    parser.add_argument('-async_meters', default=0, type=int,
                        help="1 = read loss / gradient norm of step k from pinned memory after step k+1 is enqueued: the host "
                             "stays one step ahead of the GPU (+5-8 %% throughput); the printed meters lag by one step")
    args = parser.parse_args()

    th.manual_seed(args.seed)
    np.random.seed(args.seed)
    config = _common.load_config(args.config, args.data)
    config["sweep_size"] = args.sweep_size
    config["data_path"] = args.dataPath
    _common.dump_config(config)
    rank, world, local = _common.init_distributed(True)
    if not th.cuda.is_available():
        raise SystemExit("train_chain.py: the B200 build has no CPU path")
    dev = th.device("cuda", local)
    os.makedirs(args.exp_dir, exist_ok=True)
    mc, dc = config["model_config"], config["data_config"]
    if args.synthetic > 0:
        dataset = SyntheticWaveDataset(args.synthetic, mc["label_size"])
    else:
        # corpus in the reference's formats (zip of wavs + label file, -data).  With -ali_dir / -chain_dir the labels are
        # the alignment model's transition ids and the numerator graphs are built the reference's way
        # (chain_supervision.py: phones / durations -> proto-supervision -> tree-expanded, time-constrained FST); without
        # them the labels are pdf ids and synth.alignment_to_supervision_fst builds a pdf-level graph with a tolerance
        dataset = SpeechDataset(config)
    kaldi = load_kaldi_assets(args)
    if kaldi is not None and not args.den_fst and os.path.isfile(args.chain_dir + "/den.fst"):
        args.den_fst = args.chain_dir + "/den.fst"                                   # bin/train_chain.py:167
    if args.synthetic <= 0 and not args.den_fst:
        print("WARNING: no -den_fst / <chain_dir>/den.fst: training against a synthetic denominator graph")
    supervision_opts = SupervisionOptions()
    batch_transform = None
    if kaldi is not None and args.synthetic <= 0:
        def batch_transform(batch):
            """Numerator graphs from the alignments (bin/train_chain.py:262-272), built where the batch is collated:
            in the DataLoader workers when -data_loader_threads > 0.  batch["sup"][j] = (fst dict, output frames)."""
            sup = []
            for wav, lab in zip(batch["wav"], batch["label"]):
                n = min(fb.num_frames(len(wav)), len(lab))                    # label trim, data/sr_dataset.py:358-363
                sup.append(kaldi_supervision(kaldi, supervision_opts, lab[:n, 0]))
            batch["sup"] = sup
            return batch
    loader = WaveDataloader(dataset, args.batch_size, num_workers=args.data_loader_threads, distributed=world > 1,
                            balanced=True, seed=args.seed, batch_transform=batch_transform)
    feat = pipeline.FeaturePipeline(use_cmn=dc.get("use_cmn", True))
    print("Data loader set up successfully!")
    print("Number of minibatches: {}".format(len(loader)))

    model = lstm.LSTMAM(mc["feat_dim"], mc["label_size"], mc["hidden_size"], mc["num_layers"], mc["dropout"], True).to(dev)
    optimizer = th.optim.Adam(model.parameters(), lr=args.lr, amsgrad=True, fused=True)   # one launch for all parameters
    if args.seed_model:
        _common.load_model_state(model, args.seed_model)
        print("=> loaded checkpoint '{}' ".format(args.seed_model))
    if world > 1:
        pkdist.broadcast_parameters(model)
        pkdist.broadcast_optimizer_state(optimizer)
    averager = pkdist.GradAverager(list(model.parameters())) if world > 1 else None

    chain_opts = graphs.ChainTrainingOptions(leaky_hmm_coefficient=1e-4, xent_regularize=args.xent_regularize)
    if args.den_fst:                      # a real denominator graph: OpenFst binary (Kaldi's den.fst) or fstprint text
        den = graphs.DenominatorGraph.from_file(args.den_fst, mc["label_size"])
    else:
        den = graphs.DenominatorGraph(synth.make_den_fst(args.den_states, mc["label_size"], 7, seed=1234), mc["label_size"])

    model.train()
    for epoch in range(args.num_epochs):
        loader.set_epoch(epoch)
        run_train_epoch(model, optimizer, averager, feat, loader, epoch, supervision_opts, den, chain_opts, args, rank, kaldi)
        if rank == 0:
            _common.save_checkpoint(args.exp_dir + '/chain.model.' + str(epoch) + '.tar', model, optimizer)


def kaldi_supervision(kaldi, supervision_opts, trans_ids):
    """bin/train_chain.py:262-272: alignment -> phones / durations -> proto-supervision -> supervision.  When the
    constraints leave no path (Kaldi: "Supervision FST is empty") the tolerances are doubled until one exists."""
    opts = chain_supervision.SupervisionOptions(supervision_opts.left_tolerance, supervision_opts.right_tolerance,
                                                supervision_opts.frame_subsampling_factor, True)
    while True:
        fst, t_sub = chain_supervision.supervision_from_alignment(opts, kaldi["ali_tm"], kaldi["chain_tm"], kaldi["tree"],
                                                                  trans_ids)
        if fst is not None:
            return fst, t_sub
        if opts.left_tolerance > 4 * len(trans_ids):
            raise RuntimeError("no numerator path: more phone states than output frames (%d frames)" % t_sub)
        print("WARNING: empty supervision at tolerance %d, retrying with %d" % (opts.left_tolerance, 2 * opts.left_tolerance + 1))
        opts.left_tolerance = opts.right_tolerance = 2 * opts.left_tolerance + 1


def run_train_epoch(model, optimizer, averager, feat, loader, epoch, supervision_opts, den, chain_opts, args, rank, kaldi=None):
    batch_time = utils.AverageMeter('Time', ':6.3f')
    losses = utils.AverageMeter('Loss', ':.4e')
    grad_norm = utils.AverageMeter('grad_norm', ':.4e')
    progress = utils.ProgressMeter(len(loader), batch_time, losses, grad_norm, prefix="Epoch: [{}]".format(epoch))
    rtf = utils.RTFMeter()
    factor = supervision_opts.frame_subsampling_factor
    criterion = ops.ChainObjtiveFunction.apply
    lagging = None
    end = time.time()
    for i, batch in enumerate(loader):
        wav, woff, foff = feat.ex.pack(batch["wav"])
        shift = epoch % factor                      # frame_shift = -(epoch % 3); x = roll(x, shift, 1)
        n_fr = None
        if args.synthetic <= 0:                     # label trim (data/sr_dataset.py:358-363)
            n_fr = [min(int(foff[u + 1] - foff[u]), len(l)) for u, l in enumerate(batch["label"])]
        x, num_frs = feat.sequence_batch(wav, woff, foff, factor=factor, shift=shift, n_frames=n_fr)
        # numerator graphs: the reference derives them from the alignment (bin/train_chain.py:262-272);
        # synthetic stand-in, seeded per utterance
        sups = []
        for j, ids in enumerate(batch["utt_ids"]):
            t_sub = (int(num_frs[j]) - 1) // factor + 1
            if args.synthetic > 0:
                rng = np.random.default_rng(zlib.crc32(ids[0].encode()))
                sup_fst = synth.make_supervision_fst(t_sub, den.num_pdfs(), rng)
            elif kaldi is not None:
                # the label file holds the alignment model's transition ids, as in the reference (y = trans_ids);
                # the frame shift of the epoch moves the features, not the supervision (bin/train_chain.py:251-272)
                sup_fst, t_k = batch["sup"][j]                 # built by the loader's batch_transform
                assert t_k == t_sub, (t_k, t_sub)
            else:
                sup_fst = synth.alignment_to_supervision_fst(batch["label"][j][:, 0], factor, shift,
                                                             slack=args.tolerance, n_out=t_sub)
            sups.append(graphs.Supervision(sup_fst, t_sub, den.num_pdfs()))
        prediction = model(x, valid_lengths=[s.frames_per_sequence for s in sups])   # no output layer on the padding
        if args.per_utt_loss:
            loss = 0.0
            for j, sup in enumerate(sups):
                loglike_j = prediction[j, :sup.frames_per_sequence, :]
                loss += criterion(loglike_j, den, sup, chain_opts)
        else:
            loss = ops.ChainObjtiveFunction.apply_batch(prediction, den, sups, chain_opts)
        loss.backward()
        step = len(loader) * epoch + i + 1
        lr = utils.noam_decay(step, args.warmup_steps, args.lr)
        for g in optimizer.param_groups:
            g['lr'] = lr
        tot_frs = np.array(num_frs).sum()
        if args.async_meters:
            pend_loss = pipeline.PendingValue(loss)
            norm = pipeline.finish_step(model, optimizer, averager, args.max_grad_norm)
            pend = (pend_loss, pipeline.PendingValue(norm), tot_frs)
            if lagging is not None:                  # meters of the PREVIOUS step: its copies have long landed
                losses.update(lagging[0].value() / lagging[2])
                grad_norm.update(lagging[1].value())
            lagging = pend
        else:
            norm = pipeline.finish_step(model, optimizer, averager, args.max_grad_norm)
            grad_norm.update(float(norm))
            losses.update(loss.item() / tot_frs)
        rtf.update(tot_frs)
        batch_time.update(time.time() - end)
        end = time.time()
        if rank == 0 and i % args.save_freq == 0:
            _common.save_checkpoint(args.exp_dir + '/chain.model.' + str(i) + '.tar', model, optimizer)
        if rank == 0 and i % args.print_freq == 0:
            progress.print(i)
            print("iRTF {:.1f}".format(rtf.irtf), flush=True)
        if args.max_steps and i + 1 >= args.max_steps:
            break


if __name__ == '__main__':
    main()
