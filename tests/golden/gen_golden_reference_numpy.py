"""Generates tests/golden/reference_numpy.npz by RUNNING the reference's own numpy code in this container.

The reference's Keras/TensorFlow graph cannot run here, but a few functions on and next to the path are
plain numpy inside modules that merely import TensorFlow / scipy.misc / skimage at the top.  This script
stubs those imports, loads the reference modules from /root/reference (nothing is copied into the repo)
and records their outputs on seeded inputs:

  * evaluate_models/utils.py: compute_iou (the Dice overlap of this copy), compute_overlaps,
    non_max_suppression                                    -> pins oracle/postprocess.py
  * evaluate_models/test_score_dense_captions.py: the body of refine_generations (:245-283), extracted
    with ast and executed against the reference's non_max_suppression    -> pins oracle.postprocess.refine_generations
  * dense_img_cap_separate_models/utils.py: generate_pyramid_anchors, apply_box_deltas (numpy twin of
    apply_box_deltas_graph)                                -> pins oracle/proposals.py

  * evaluate_models/modified_dense_model.py: class PyramidROIAlign (+ log2_graph), and
    dense_img_cap_separate_models/modified_dense_model.py: class ProposalLayer (+ apply_box_deltas_graph,
    clip_boxes_graph, utils.batch_slice), extracted with ast and EXECUTED over the numpy stand-in for TensorFlow of
    tf_numpy_shim.py (primitives = the oracle's restatements; the glue around them is what is pinned)
                                                           -> pins oracle.pyramid_roi_align_literal / proposal_layer
  * dense_img_cap_separate_models/preprocess.py: load_corpus, encode_caption (nltk stubbed with the product's
    regular-expression tokenizer: the token FILTERING is what is pinned), and, extracted with ast from
    text_generation_model.py / text_generation_model_v2.py, data_generator (:330-372), load_sequences
    (:128-137) and the v2 script's own data_generator (:169-205) run on a small fake dataset
                                                           -> pins image_captioning_b200/data.py

Scores are made distinct: numpy's default argsort is not stable, so tie order is not a property of the
reference.  Run from the repo root (only where /root/reference exists):
    python tests/golden/gen_golden_reference_numpy.py
"""
import ast
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_numpy_shim as shim  # noqa: E402

for name in ("scipy.misc", "skimage", "skimage.color", "skimage.io"):
    if name not in sys.modules or name == "scipy.misc":
        _stub(name)
tf = shim.make_tf()                                         # numpy stand-in for the TF graph ops (tf_numpy_shim.py)
sys.modules["tensorflow"] = tf


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ev = _load(os.path.join(REF, "evaluate_models", "utils.py"), "ref_eval_utils")
dn = _load(os.path.join(REF, "dense_img_cap_separate_models", "utils.py"), "ref_dense_utils")

# refine_generations is a method of a class in a module with heavy imports: take the function node only
src = open(os.path.join(REF, "evaluate_models", "test_score_dense_captions.py")).read()
fn = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "refine_generations"][0]
ns = {"np": np, "non_max_suppression": ev.non_max_suppression}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "refine_generations", "exec"), ns)
ref_refine = ns["refine_generations"]

rng = np.random.default_rng(20261019)
out = {}

# ---- post-processing ------------------------------------------------------------------------------
cases = []
for n, thr in ((1, 0.5), (17, 0.5), (200, 0.5), (200, 0.3), (300, 0.7)):
    c = rng.uniform(0.15, 0.85, (n, 2))
    half = rng.uniform(0.01, 0.25, (n, 2))
    boxes = np.concatenate([c - half, c + half], 1).astype(np.float32)
    P, V = 5, 12
    probs = rng.dirichlet(np.ones(V) * 0.3, (n, P)).astype(np.float32)
    scores = np.sum(np.log(np.max(probs, axis=2)), axis=1)
    assert len(np.unique(scores)) == n
    keep = ev.non_max_suppression(boxes, scores, thr)

    class Cfg:
        DETECTION_NMS_THRESHOLD = thr
        DETECTION_MAX_INSTANCES = 100 if n != 300 else 20
    rois, caps = ref_refine(None, boxes, probs, None, Cfg)
    cases.append((boxes, probs, scores, keep, rois, caps, thr, Cfg.DETECTION_MAX_INSTANCES))
for i, (boxes, probs, scores, keep, rois, caps, thr, mx) in enumerate(cases):
    out["pp%d_boxes" % i], out["pp%d_probs" % i], out["pp%d_scores" % i] = boxes, probs, scores
    out["pp%d_nms_keep" % i], out["pp%d_refined_rois" % i] = keep, rois
    out["pp%d_refined_first_probs" % i] = caps[:, 0, :]
    out["pp%d_params" % i] = np.array([thr, mx], np.float64)
b1 = cases[1][0]
area = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
out["overlaps_17"] = ev.compute_overlaps(b1, b1[:5])
out["iou_17_row0"] = ev.compute_iou(b1[0], b1, area[0], area)

# ---- anchors --------------------------------------------------------------------------------------
scales, ratios, strides = (32, 64, 128, 256, 512), [0.5, 1, 2], [4, 8, 16, 32, 64]
small = np.array([[int(np.ceil(128 / s)), int(np.ceil(128 / s))] for s in strides])
full = np.array([[int(np.ceil(1024 / s)), int(np.ceil(1024 / s))] for s in strides])
out["anchors_128"] = dn.generate_pyramid_anchors(scales, ratios, small, strides, 1)
a_full = dn.generate_pyramid_anchors(scales, ratios, full, strides, 1)
out["anchors_1024_shape"] = np.array(a_full.shape)
out["anchors_1024_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a_full.astype(np.float32)).tobytes()).digest(), np.uint8)
out["anchors_128_stride2"] = dn.generate_pyramid_anchors(scales[:2], ratios, small[:2], strides[:2], 2)

# ---- box deltas (numpy twin of apply_box_deltas_graph) ------------------------------------------
anc = out["anchors_128"].astype(np.float32)
deltas = (rng.standard_normal(anc.shape) * np.array([0.1, 0.1, 0.2, 0.2]) * 3).astype(np.float32)
out["deltas_128"] = deltas
out["refined_128"] = dn.apply_box_deltas(anc, deltas)
assert out["refined_128"].dtype == np.float32

# ---- data formats -----------------------------------------------------------------------------------
import importlib
data_mod = importlib.import_module("image_captioning_b200.data")
nltk = _stub("nltk"); tok = _stub("nltk.tokenize"); tok.word_tokenize = data_mod.tokenize; nltk.tokenize = tok
np.float = float                                           # preprocess.py predates numpy 1.20
pre = _load(os.path.join(REF, "dense_img_cap_separate_models", "preprocess.py"), "ref_preprocess")
vocab_tokens = ["a", "man", "dog", "red", "on", "the", "grass", "frisbee", "with", "'s", "two", "playing"]
emb = {t: rng.standard_normal(6) for t in vocab_tokens}
np.random.seed(11)
w2i, i2w, mat = pre.load_corpus(vocab_tokens, emb, 6)
out["corpus_matrix"] = mat
out["corpus_words"] = np.array([i2w[i] for i in range(len(i2w))])
captions = ["A man's dog on the grass.", "Two dogs playing with a red frisbee", "zebra", "The man, the dog; the GRASS!"]
out["captions_text"] = np.array(captions)
enc = [pre.encode_caption(c, w2i) for c in captions]
out["captions_encoded"] = np.array([",".join(str(int(v)) for v in e) for e in enc])


def _extract(path, name):
    tree = ast.parse(open(path).read())
    return [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name][0]


P, V = 6, len(w2i)
caps_by_image = {7: np.array([[1, 4, 5, 2, 0, 0], [1, 3, 2, 0, 0, 0]], np.float32), 9: np.array([[1, 8, 9, 10, 11, 2]], np.float32)}
feats_by_image = {k: rng.standard_normal((v.shape[0], 2, 2, 3)).astype(np.float32) for k, v in caps_by_image.items()}


class FakeDataset:
    _image_ids = [7, 9]
    rois = [(k, i, caps_by_image[k][i]) for k in (7, 9) for i in range(caps_by_image[k].shape[0])]

    def load_captions_and_rois(self, image_id):
        c = caps_by_image[image_id].astype(np.int64)
        onehot = np.zeros(c.shape + (V,))
        np.put_along_axis(onehot, c[..., None], 1.0, -1)
        return None, onehot


class Cfg2:
    VOCABULARY_SIZE = V


ns = {"np": np, "generate_features": lambda dataset, image_id, model: feats_by_image[image_id]}
exec(compile(ast.Module(body=[_extract(os.path.join(REF, "dense_img_cap_separate_models", "text_generation_model.py"), "data_generator")],
                        type_ignores=[]), "data_generator", "exec"), ns)
gen = ns["data_generator"](FakeDataset(), None, Cfg2, 2, shuffle=False)
batches = [next(gen) for _ in range(3)]                      # 3 RoIs, batch 2: wraps around
out["gen_features"] = np.stack([b[0][0] for b in batches])
out["gen_words"] = np.stack([b[0][1] for b in batches])
out["gen_targets"] = np.stack([b[1] for b in batches])
out["gen_caps_7"], out["gen_caps_9"] = caps_by_image[7], caps_by_image[9]
out["gen_feats_7"], out["gen_feats_9"] = feats_by_image[7], feats_by_image[9]
ns2 = {"np": np, "tqdm": lambda x: x}
exec(compile(ast.Module(body=[_extract(os.path.join(REF, "dense_img_cap_separate_models", "text_generation_model_v2.py"), "load_sequences")],
                        type_ignores=[]), "load_sequences", "exec"), ns2)
seqs = ns2["load_sequences"](FakeDataset())
out["v2_sequences"] = np.array(["%d|%d|%s|%d" % (a, b, ",".join(str(int(t)) for t in c), int(d)) for a, b, c, d in seqs])

# the v2 script's own data_generator (text_generation_model_v2.py:169-205) on those sequences: 18 sequences, batches of 4,
# five batches = one wrap-around; keras' pad_sequences is the oracle's restatement of its defaults (pre-padding)
from oracle import decoder as _dec_pad


class FakeSeqDataset(FakeDataset):
    sequences = seqs


class Cfg3:
    VOCABULARY_SIZE = V
    PADDING_SIZE = 4                                           # shorter than the longest prefix (5): pre-truncation too


ns2b = {"np": np, "generate_features": lambda dataset, image_id, model: feats_by_image[image_id],
        "pad_sequences": lambda sequences, maxlen: _dec_pad.pad_sequences_pre(sequences, maxlen)}
exec(compile(ast.Module(body=[_extract(os.path.join(REF, "dense_img_cap_separate_models", "text_generation_model_v2.py"), "data_generator")],
                        type_ignores=[]), "data_generator_v2", "exec"), ns2b)
gen2 = ns2b["data_generator"](FakeSeqDataset(), None, Cfg3, 4, shuffle=False)
batches2 = [next(gen2) for _ in range(5)]
out["v2gen_features"] = np.stack([b[0][0] for b in batches2])
out["v2gen_words"] = np.stack([b[0][1] for b in batches2])
out["v2gen_next"] = np.stack([b[1] for b in batches2])

# ---- beam search control flow -----------------------------------------------------------------------
# gen_captions ("image captioning/test.py":23-64) is plain Python around model.predict: run it with a stand-in model
# whose predict() is the oracle's v1 word model (float64 copy of the fp32 probabilities, so that the score sums are
# float64 under numpy 1.x and 2.x alike) -> pins the candidate / pooling / sorting / scoring logic of oracle.beam_v1
import contextlib
import io
from oracle import decoder as dec
from image_captioning_b200 import synth
BV, BP, BK, BR = 40, 6, 3, 5
wb = synth.synth_weights_v1(np.random.default_rng(77), V=BV, E=6, F=16, U=8, pool=2, C=4, trained_like=False)
fb = dec.head(np.random.default_rng(78).standard_normal((BR, 2, 2, 4)).astype(np.float32), wb)


class FakeModel:
    def predict(self, inputs):
        f, cap = np.asarray(inputs[0], np.float32), np.asarray(inputs[1])[0]
        st = dec.V1State(1, 8)
        for t in cap[cap != 0]:                                # pre-padding zeros are masked steps: state untouched
            p = dec.v1_step(f, np.array([t]), st, wb)
        return p.astype(np.float64)


class FakeTokenizer:
    def texts_to_sequences(self, texts):
        return [[1]]


ns3 = {"np": np, "pad_sequences": lambda sequences, maxlen, padding: dec.pad_sequences_pre(sequences, maxlen),
       "make_caption_human_readable": lambda caption, index_to_word: ""}
exec(compile(ast.Module(body=[_extract(os.path.join(REF, "image captioning", "test.py"), "gen_captions")], type_ignores=[]),
             "gen_captions", "exec"), ns3)
btok, bsc = np.zeros((BR, BK, BP), np.int32), np.zeros((BR, BK))
for r in range(BR):
    with contextlib.redirect_stdout(io.StringIO()):
        res = ns3["gen_captions"]({"max_caption_length": BP}, FakeModel(), fb[r], FakeTokenizer(), BK, None)
    for j, (cap, sc) in enumerate(res):
        btok[r, j], bsc[r, j] = cap, sc
out["beam_ref_tokens"], out["beam_ref_scores"] = btok, bsc
out["beam_params"] = np.array([BV, BP, BK, BR])

# ---- v2 greedy loop control flow ------------------------------------------------------------------------
# the per-RoI loop of evaluate_models/test_score_dense_captions.py:214-225 (start word zeros(V) -> id 0, P-1 predicts over
# the pre-padded argmax history, probabilities collected), extracted as an AST statement and run with a stand-in
# model.predict (the oracle's v2 inject model) -> pins the loop logic of oracle.greedy_v2
score_src = open(os.path.join(REF, "evaluate_models", "test_score_dense_captions.py")).read()
line_no = 1 + [i for i, l in enumerate(score_src.split("\n")) if l.strip() == "for j in range(img_boxes.shape[0]):"][0]
loop = [n for n in ast.walk(ast.parse(score_src)) if isinstance(n, ast.For) and n.lineno == line_no][0]
GV, GP, GR = 30, 6, 4
wg = synth.synth_weights_v2(np.random.default_rng(79), V=GV, E=6, F=16, units=8, pool=2, C=4, trained_like=False)
featg = np.random.default_rng(80).standard_normal((GR, 2, 2, 4)).astype(np.float32)


class V2Model:
    def predict(self, inputs):
        return dec.v2_inject_predict(np.asarray(inputs[0], np.float32), np.asarray(inputs[1]), wg)


class Self:
    class config:
        VOCABULARY_SIZE, PADDING_SIZE = GV, GP
    model = V2Model()


ns4 = {"np": np, "self": Self, "img_boxes": np.zeros((GR, 4)), "img_features": featg, "caps": [],
       "pad_sequences": lambda seqs, maxlen: dec.pad_sequences_pre(seqs, maxlen)}
exec(compile(ast.Module(body=[loop], type_ignores=[]), "v2_greedy_loop", "exec"), ns4)
out["v2_loop_probs"] = np.array(ns4["caps"])                   # [R, P-1, V]
out["v2_loop_params"] = np.array([GV, GP, GR])
# the same loop as evaluate_models/eval_text_generation_model_v2.py:176-186 runs it: started from the GT first word
eval_src = open(os.path.join(REF, "evaluate_models", "eval_text_generation_model_v2.py")).read()
line_no = 1 + [i for i, l in enumerate(eval_src.split("\n")) if l.strip() == "for j in range(captions.shape[0]):"][0]
loop2 = [n for n in ast.walk(ast.parse(eval_src)) if isinstance(n, ast.For) and n.lineno == line_no][0]
start_ids = np.array([3, 7, 1, 12])
gt_onehot = np.zeros((GR, GP, GV))
gt_onehot[np.arange(GR), 0, start_ids] = 1.0
words_seen = []
ns4b = {"np": np, "features": featg, "captions": gt_onehot, "model": V2Model(), "predictions": [],
        "config": Self.config, "id_to_word": None, "pad_sequences": lambda seqs, maxlen: dec.pad_sequences_pre(seqs, maxlen),
        "decode_word": lambda p, id_to_word: words_seen.append(int(np.argmax(p))) or str(int(np.argmax(p)))}
exec(compile(ast.Module(body=[loop2], type_ignores=[]), "v2_eval_loop", "exec"), ns4b)
out["v2_eval_start"] = start_ids
out["v2_eval_predicted"] = np.array([[int(t) for t in d["p"].split(" ")] for d in ns4b["predictions"]])   # [R, P] ids incl. the start word

# ---- the two TF layers on the path, executed over the numpy stand-in ----------------------------------------
def _extract_defs(path, names):
    tree = ast.parse(open(path).read())
    return [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]


KE = types.SimpleNamespace(Layer=shim.Layer)
ns5 = {"tf": tf, "np": np, "KE": KE}
exec(compile(ast.Module(body=_extract_defs(os.path.join(REF, "evaluate_models", "modified_dense_model.py"),
                                           ("log2_graph", "PyramidROIAlign")), type_ignores=[]), "PyramidROIAlign", "exec"), ns5)
gr = np.load(os.path.join(ROOT, "tests", "golden", "roi_align_small.npz"))
layer = ns5["PyramidROIAlign"]([7, 7], tuple(int(v) for v in gr["image_shape"]))
pooled_ref = np.asarray(layer.call([shim.tensor(gr["boxes"])] + [shim.tensor(gr[k]) for k in ("p2", "p3", "p4", "p5")]))
assert pooled_ref.shape == gr["pooled"].shape and pooled_ref.dtype == np.float32
out["roi_align_layer_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pooled_ref).tobytes()).digest(), np.uint8)
out["roi_align_layer_equals_oracle_golden"] = np.array(np.array_equal(pooled_ref.view(np.uint32), gr["pooled"].view(np.uint32)))
assert layer.compute_output_shape([(2, 48, 4), (2, 32, 32, 8)]) == (2, 48, 7, 7, 8)

ns6 = {"tf": tf, "np": np, "KE": KE, "utils": dn}
exec(compile(ast.Module(body=_extract_defs(os.path.join(REF, "dense_img_cap_separate_models", "modified_dense_model.py"),
                                           ("apply_box_deltas_graph", "clip_boxes_graph", "ProposalLayer")), type_ignores=[]),
             "ProposalLayer", "exec"), ns6)


class PCfg:
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    IMAGES_PER_GPU = 2
    IMAGE_SHAPE = np.array([128, 128, 3])


anchors_p = out["anchors_128"][1800:3400]              # P2 tail + all of P3: 1600 anchors keep the fixture small
A = anchors_p.shape[0]
fgp = (1.0 / (1.0 + np.exp(-(rng.standard_normal((2, A)) * 2.5 - 3.0)))).astype(np.float32)
fgp = np.round(fgp, 3)                                       # score ties
probs_p = np.stack([1 - fgp, fgp], -1).astype(np.float32)
bbox_p = (rng.standard_normal((2, A, 4)) * 1.5).astype(np.float32)
bbox_p[0, ::9] = [0.0, 0.0, -80.0, -80.0]                    # degenerate boxes
bbox_p[1, ::7, :2] = 40.0                                    # pushed out of the image, clipped to a corner
player = ns6["ProposalLayer"](proposal_count=1500, nms_threshold=0.7, anchors=anchors_p, config=PCfg)   # > survivors: padding
prop_ref = np.asarray(player.call([shim.tensor(probs_p), shim.tensor(bbox_p)]))
assert prop_ref.shape == (2, 1500, 4) and prop_ref.dtype == np.float32 and player.compute_output_shape(None) == (None, 1500, 4)
out["proposal_probs"], out["proposal_bbox"], out["proposal_ref"] = probs_p, bbox_p, prop_ref
out["proposal_anchor_slice"] = np.array([1800, 3400])

# ---- v1 caption models: the reference's WIRING executed over eager stand-ins for the Keras layers ---------------
# word_generation_model, ROICaptionInferenceLayer (greedy decoding), build_roi_caption_model_training (teacher forcing)
# and roi_caption_loss are run from the reference source; layer numerics = the oracle's restatements.
KL, KM, K = shim.make_keras(tf)
tg_path = os.path.join(REF, "dense_img_cap_separate_models", "text_generation_model.py")
ns7 = {"tf": tf, "np": np, "KL": KL, "KM": KM, "K": K}
exec(compile(ast.Module(body=_extract_defs(tg_path, ("word_generation_model", "build_roi_caption_model_training",
                                                     "ROICaptionInferenceLayer", "roi_caption_loss")), type_ignores=[]),
             "text_generation_model", "exec"), ns7)
CV, CP, CB, CU, CF = 30, 5, 3, 8, 16
wc = synth.synth_weights_v1(np.random.default_rng(81), V=CV, E=6, F=CF, U=CU, pool=2, C=4, trained_like=False)
shim.WEIGHTS.update(wc)
fc = dec.head(np.random.default_rng(82).standard_normal((CB, 2, 2, 4)).astype(np.float32), wc)
gtc = synth.synth_captions(np.random.default_rng(83), CB, CP, CV)
gtc[1, 2] = 0                                                  # a masked step in the middle of a caption


class CCfg:
    BATCH_SIZE, PADDING_SIZE, VOCABULARY_SIZE, EMBEDDING_SIZE = CB, CP, CV, 6
    EMBEDDING_WEIGHTS = wc["imgcap_embedding_layer/embeddings"]


ref_wgm = ns7["word_generation_model"]


def recallable_word_model(features_input, lstm_units, config):
    """The reference builds the word model once and calls it many times; the eager stand-in re-runs the reference's
    builder on every call with the call's input fed to its KL.Input."""
    def call(x):
        shim.FEEDS["word_model_input"] = np.asarray(x)
        return ref_wgm(features_input, lstm_units, config).outputs
    return call


ns7["word_generation_model"] = recallable_word_model
word_model = recallable_word_model([CF + CP], CU, CCfg)
greedy_probs = np.asarray(ns7["ROICaptionInferenceLayer"](word_model, CCfg).call(shim.tensor(fc)))
shim.FEEDS["input_imgcap_caption_features"] = np.concatenate([fc[:, None, :], gtc[:, None, :]], -1)     # [B, 1, F + P]
train_probs = np.asarray(ns7["build_roi_caption_model_training"]([1, CF + CP], CU, CCfg).outputs)
ids_c = dec.targets_from_captions(gtc)
onehot_c = np.zeros((CB, CP, CV), np.float32)
np.put_along_axis(onehot_c, ids_c[..., None].astype(np.int64), 1.0, -1)
onehot_c[2, 3:] = 0.0                                          # positions without a target are excluded from the mean
out["v1_greedy_probs"], out["v1_train_probs"], out["v1_gt"] = greedy_probs, train_probs, gtc
out["v1_loss"] = np.asarray(ns7["roi_caption_loss"](shim.tensor(onehot_c), shim.tensor(train_probs)), np.float32)
out["v1_loss_all_masked"] = np.asarray(ns7["roi_caption_loss"](shim.tensor(onehot_c * 0), shim.tensor(train_probs)), np.float32)
out["v1_params"] = np.array([CV, CP, CB, CU, CF])

# the whole models: build_lstm_model (RoI head -> TimeDistributed caption layer / training graph) and the v2 build_model
ns7["BatchNorm"] = ns8_bn = None
bn_ns = {"KL": KL}
exec(compile(ast.Module(body=_extract_defs(os.path.join(REF, "dense_img_cap_separate_models", "modified_dense_model.py"), ("BatchNorm",)),
                        type_ignores=[]), "BatchNorm", "exec"), bn_ns)
ns7["BatchNorm"] = bn_ns["BatchNorm"]
exec(compile(ast.Module(body=_extract_defs(tg_path, ("build_lstm_model",)), type_ignores=[]), "build_lstm_model", "exec"), ns7)
ref_train_builder = ns7["build_roi_caption_model_training"]


def recallable_training_model(features_input, lstm_units, config):
    m = KM.Model(None, None)

    def recall(x):
        shim.FEEDS["input_imgcap_caption_features"] = np.asarray(x)
        return ref_train_builder(features_input, lstm_units, config).outputs
    m.recall = recall
    return m


ns7["build_roi_caption_model_training"] = recallable_training_model
CCfg.POOL_SIZE = 2
wc2 = synth.synth_weights_v1(np.random.default_rng(84), V=CV, E=6, F=1024, U=CU, pool=2, C=4, trained_like=False)   # the head's 1024
shim.WEIGHTS.clear()                                                                                                 # filters are hardcoded
shim.WEIGHTS.update(wc2)
CCfg.EMBEDDING_WEIGHTS = wc2["imgcap_embedding_layer/embeddings"]
featc = np.random.default_rng(82).standard_normal((CB, 2, 2, 4)).astype(np.float32)
shim.FEEDS["input_imgcap_lstm_features"] = featc
shim.FEEDS["input_imgcap_lstm_gt_captions"] = gtc
out["v1_model_inference"] = np.asarray(ns7["build_lstm_model"]([2, 2, 4], CCfg, CU, "inference").outputs)
out["v1_model_training"] = np.asarray(ns7["build_lstm_model"]([2, 2, 4], CCfg, CU, "training").outputs)

v2_path = os.path.join(REF, "dense_img_cap_separate_models", "text_generation_model_v2.py")
ns8 = {"tf": tf, "np": np, "KL": KL, "KM": KM, "BatchNorm": bn_ns["BatchNorm"]}
exec(compile(ast.Module(body=_extract_defs(v2_path, ("build_model",)), type_ignores=[]), "build_model_v2", "exec"), ns8)
wg2 = synth.synth_weights_v2(np.random.default_rng(85), V=GV, E=6, F=1024, units=8, pool=2, C=4, trained_like=False)
shim.WEIGHTS.clear()
shim.WEIGHTS.update(wg2)
KL._counters.clear()


class V2Cfg:
    POOL_SIZE, VOCABULARY_SIZE, EMBEDDING_SIZE = 2, GV, 6
    EMBEDDING_WEIGHTS = wg2["imgcap_embedding_layer/embeddings"]


words_v2 = dec.pad_sequences_pre([[0], [3, 7], [5, 0, 9, 2], [1, 2, 3, 4, 5, 6, 7]], GP)
shim.FEEDS["imgcap_features"], shim.FEEDS["__unnamed__"] = featg, [words_v2.astype(np.float32)]
out["v2_model_probs"] = np.asarray(ns8["build_model"]([2, 2, 4], [GP], V2Cfg, 8, inject=True).outputs)
out["v2_model_words"] = words_v2

path = os.path.join(ROOT, "tests", "golden", "reference_numpy.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
