"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Pulls the reference's model classes out of /root/reference *without importing the
scripts*: the scripts import un-installed packages at the top
(`editnet.py:15-16`) and open data files / start training at module scope
(`editnet.py:743-847`).  We `ast.parse` the file, keep only the `ClassDef` nodes we
ask for, and `exec` them in a namespace that pre-binds what their bodies use.
Nothing is copied into this repository; the classes live only in memory.

Only usable where /root/reference exists (the authoring container).  The GPU box
has no reference tree: tests that need it skip there, and the committed
`tests/golden/*.npz` fixtures (written by `oracle/make_golden.py` from these very
classes) carry the reference's outputs instead.
"""
import ast
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import PackedSequence, pack_padded_sequence, pad_packed_sequence

REFERENCE_ROOT = os.environ.get("SET_REFERENCE_ROOT", "/root/reference")

EDITNET_CLASSES = ("LSTMCellC", "CopyLSTMCellC", "EmbeddingC", "CaptionEncoderC",
                   "CaptionAttentionC", "SelectC", "VisualAttentionC", "DecoderC")
DCNET_CLASSES = ("Embedding", "CaptionEncoder", "CaptionAttention", "DAE")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "editnet.py"))


def extract_classes(rel_path, names, extra=(), device="cpu"):
    """exec the named top-level classes of `rel_path`; returns the namespace dict."""
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r") as f:
        src = f.read()
    try:
        tree = ast.parse(src)
    except SyntaxError:
        # dcnet_with_mse.py:346-349 has an IndentationError; callers slice by lines.
        raise
    wanted = set(names) | set(extra)
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in wanted]
    missing = wanted - {n.name for n in body}
    if missing:
        raise KeyError("classes not found in %s: %s" % (rel_path, sorted(missing)))
    mod = ast.Module(body=body, type_ignores=[])
    ns = {
        "torch": torch, "nn": nn, "F": F, "np": np, "math": math,
        "device": torch.device(device),
        "pack_padded_sequence": pack_padded_sequence,
        "pad_packed_sequence": pad_packed_sequence,
        "PackedSequence": PackedSequence,
    }
    exec(compile(mod, path, "exec"), ns)
    return ns


def editnet_xe_classes():
    return extract_classes("editnet.py", EDITNET_CLASSES)


def editnet_rl_classes():
    return extract_classes("editnet_rl.py", EDITNET_CLASSES, extra=("RewardCriterion",))


def editnet_adaptive_classes():
    return extract_classes("adaptive_features/editnet_adaptive.py", EDITNET_CLASSES)


def dcnet_xe_classes():
    return extract_classes("dcnet.py", DCNET_CLASSES)


def dcnet_rl_classes():
    return extract_classes("dcnet_rl.py", DCNET_CLASSES)


def eval_full_search(device="cpu"):
    """The reference's ensemble search, `evaluate_full` of eval/eval xe/eval_full.py:88-215, as a callable
    `(loader, dae_ar, decoder, beam_size, epoch, word_map) -> results` (list of {"image_id", "caption"}).

    Two edits, both outside the arithmetic under test: the statement `prev_word_inds = top_k_words / vocab_size`
    (:162) becomes `//` (true division crashes on torch >= 1.5), and the function is cut after its
    per-image loop (the tail needs the COCO tool-chain: Java tokenizer, METEOR)."""
    path = os.path.join(REFERENCE_ROOT, "eval", "eval xe", "eval_full.py")
    with open(path, "r") as f:
        tree = ast.parse(f.read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "evaluate_full"][0]
    loop_at = max(i for i, n in enumerate(fn.body) if isinstance(n, ast.For))
    fn.body = fn.body[:loop_at + 1] + [ast.Return(value=ast.Name(id="results", ctx=ast.Load()))]

    class FloorDiv(ast.NodeTransformer):
        def visit_Assign(self, node):
            self.generic_visit(node)
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id == "prev_word_inds" and isinstance(node.value, ast.BinOp) \
                    and isinstance(node.value.op, ast.Div):
                node.value.op = ast.FloorDiv()
            return node

    fn = ast.fix_missing_locations(FloorDiv().visit(fn))
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "device": torch.device(device), "tqdm": lambda it, **kw: it}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["evaluate_full"]


def editnet_evaluate_search(device="cpu"):
    """The reference's own beam search, `evaluate` of editnet.py:595-719, as a callable
    `(loader, decoder, beam_size, epoch, vocab_size, word_map) -> results`; same two edits as eval_full_search (the `/`
    of :666 becomes `//`, the COCO scoring tail behind the per-image loop is cut)."""
    path = os.path.join(REFERENCE_ROOT, "editnet.py")
    with open(path, "r") as f:
        tree = ast.parse(f.read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "evaluate"][0]
    loop_at = max(i for i, n in enumerate(fn.body) if isinstance(n, ast.For))
    fn.body = fn.body[:loop_at + 1] + [ast.Return(value=ast.Name(id="results", ctx=ast.Load()))]

    class FloorDiv(ast.NodeTransformer):
        def visit_Assign(self, node):
            self.generic_visit(node)
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id == "prev_word_inds" and isinstance(node.value, ast.BinOp) \
                    and isinstance(node.value.op, ast.Div):
                node.value.op = ast.FloorDiv()
            return node

    fn = ast.fix_missing_locations(FloorDiv().visit(fn))
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "device": torch.device(device), "tqdm": lambda it, **kw: it}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["evaluate"]


def eval_class_modules():
    """the class-only copies the eval scripts import (`from dae import *; from editnet import *`, eval_full.py:18-19)"""
    e = extract_classes(os.path.join("eval", "eval xe", "editnet.py"), EDITNET_CLASSES)
    d = extract_classes(os.path.join("eval", "eval xe", "dae.py"), DCNET_CLASSES)
    return e, d


class DropoutScript:
    """Feeds pre-drawn keep-masks to every dropout call of an exec'd reference
    module, in call order, so that a train-mode reference run is reproducible and
    comparable with an implementation that takes explicit masks.

    `masks` is a list of float tensors holding 0 / 1 keep flags; the call scales by
    1/(1-p) exactly as `nn.Dropout` does.
    """

    def __init__(self, masks):
        self.masks = list(masks)
        self.calls = 0
        self._orig = None

    def __enter__(self):
        script = self
        self._orig = nn.Dropout.forward

        def forward(mod, x):
            if not mod.training:
                return x
            m = script.masks[script.calls]
            script.calls += 1
            assert m.shape == x.shape, (script.calls - 1, tuple(m.shape), tuple(x.shape))
            return x * m.to(x.dtype) * (1.0 / (1.0 - mod.p))

        nn.Dropout.forward = forward
        return self

    def __exit__(self, *exc):
        nn.Dropout.forward = self._orig
        return False
