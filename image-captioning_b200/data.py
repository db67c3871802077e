"""Data formats on either side of the path: what the reference feeds into and
reads out of the caption models.  Host-side Python, like the reference's own glue -- nothing here computes
on the path.

  * vocabulary / embedding tables      dense_img_cap_separate_models/preprocess.py:8-41 (load_corpus,
                                       load_embeddings) and the pickles of text_generation_model.py:389-414
  * caption encoding and framing       preprocess.py:59-84 (encode_caption / encode_word) and
                                       VisualGenomeDataset.load_captions_and_rois, text_generation_model.py:101-116
                                       (<start>=1 ... <end>=2, cut to PADDING_SIZE, zero post-padding, float32)
  * Visual Genome region_descriptions  text_generation_model.py:57-80 (rois as [y, x, y+h, x+w] pixels + phrases)
  * v1 training batches                data_generator, text_generation_model.py:330-372
                                       ([features, input_words], targets = input shifted left ++ [0])
  * v2 training sequences              load_sequences, text_generation_model_v2.py:128-137
  * Keras HDF5 weight files            model.load_weights(path, by_name=True) (text_generation_model.py:468, 484;
                                       modified_dense_model.py:1583-1619): layer_names / weight_names attributes.
                                       Needs h5py, which this image does not have: the group walker below works on
                                       any h5py-like mapping and is tested with a stand-in; `load_keras_h5_weights`
                                       raises ImportError with that explanation when h5py is missing.

Deviation: the reference tokenises with nltk.word_tokenize (nltk is not installed here).  `tokenize` below is a small
regular-expression tokenizer that agrees with it on plain Visual Genome phrases (lower-case words, numbers,
apostrophe clitics, punctuation split off); pass `tokenizer=nltk.word_tokenize` for the reference's exact behaviour.
"""
import json
import pickle
import re

import numpy as np

UNK, START, END = 0, 1, 2
_TOKEN = re.compile(r"n't|'(?:s|re|ve|ll|d|m)\b|[a-z0-9]+(?:[-.][a-z0-9]+)*(?=n't)|[a-z0-9]+(?:[-.][a-z0-9]+)*|[^\sa-z0-9]")


def tokenize(caption):
    """Lower-cased word / clitic / punctuation tokens (see the module note on nltk)."""
    return _TOKEN.findall(caption.lower())


def load_embeddings(file_name):
    """GloVe text file -> {word: float64 vector} (preprocess.py:31-41; words lower-cased)."""
    out = {}
    with open(file_name, "r", encoding="utf-8") as doc:
        for line in doc:
            parts = line.rstrip("\n").lower().split(" ")
            if len(parts) > 1:
                out[parts[0]] = np.array(parts[1:], dtype=np.float64)
    return out


def load_corpus(tokens, embeddings, embeddings_dim, rand=None):
    """Vocabulary tables (preprocess.py:8-28): id 0 '<unk>' (zero vector), 1 '<start>' and 2 '<end>' drawn uniformly
    from [-0.5, 0.5), then the tokens in the given order.  `rand(n)` supplies the uniform [0,1) draws (default:
    numpy's global generator, as the reference)."""
    rand = np.random.rand if rand is None else rand
    tokens = list(tokens)
    matrix = np.zeros((len(tokens) + 3, embeddings_dim))
    id_to_word = {UNK: "<unk>", START: "<start>", END: "<end>"}
    matrix[START] = rand(embeddings_dim) - 0.5
    matrix[END] = rand(embeddings_dim) - 0.5
    for i, tok in enumerate(tokens):
        id_to_word[i + 3] = tok
        matrix[i + 3] = embeddings[tok]
    word_to_id = {w: i for i, w in id_to_word.items()}
    return word_to_id, id_to_word, matrix


def load_vocabulary(id_to_word_file, word_to_id_file, embedding_matrix_file):
    """The three pickles the reference caches next to the dataset (text_generation_model.py:389-414)."""
    out = []
    for path in (word_to_id_file, id_to_word_file, embedding_matrix_file):
        with open(path, "rb") as f:
            out.append(pickle.load(f))
    return out[0], out[1], np.asarray(out[2])


def encode_word(word, word_to_id):
    return word_to_id.get(word, UNK)


def encode_caption(caption, word_to_id, tokenizer=tokenize):
    """Caption string -> int array of word ids; out-of-vocabulary tokens are DROPPED (preprocess.py:59-67)."""
    ids = [encode_word(t, word_to_id) for t in tokenizer(caption.lower())]
    return np.array([i for i in ids if i != UNK], dtype=np.int64)


def frame_caption(ids, padding_size):
    """<start> ids[:P-2] <end>, zero post-padded to P, float32 (text_generation_model.py:109-113)."""
    ids = np.asarray(ids).ravel()[:max(padding_size - 2, 0)]
    out = np.zeros((padding_size,), np.float32)
    out[0] = START
    out[1:1 + ids.size] = ids
    out[1 + ids.size] = END
    return out


def decode_caption(ids, id_to_word, stop=None):
    """ids -> ' '.join(words), optionally cut at the first `stop` word (e.g. '<end>')."""
    words = []
    for i in np.asarray(ids).ravel():
        w = id_to_word[int(i)]
        if stop is not None and w == stop:
            break
        words.append(w)
    return " ".join(words)


def generate_predictions(model, dataset, features_fn, config, id_to_word, separator=" <end>", cache_path=None, progress=None):
    """The evaluation loop around the path (evaluate_models/eval_text_generation_model.py:129-153): for every image of
    ``dataset`` the RoI features (``features_fn(image_id)`` -- ROIAlign of the image's RoIs), ``model.predict(features,
    batch_size=config.BATCH_SIZE)``, each caption decoded to words and cut at the first ``separator``, paired with the
    lower-cased ground-truth phrase: ``[{'p': predicted, 'r': real}, ...]``.  ``predict`` may return token ids [N, P] or
    per-word distributions [N, P, V] (arg-max taken, as decode_word does).  ``cache_path``: the reference's pickle cache
    (protocol 2) -- read when it exists, written otherwise.  ``progress`` wraps the image loop (tqdm in the reference)."""
    import os
    import pickle
    if cache_path and os.path.exists(cache_path):
        with open(cache_path, "rb") as f:
            return pickle.load(f)
    predictions = []
    ids = dataset.image_ids
    for image_id in (progress(ids) if progress else ids):
        features = features_fn(image_id)
        _, captions = dataset.load_original_captions_and_rois(image_id)
        result = np.asarray(model.predict(features, batch_size=config.BATCH_SIZE))
        if result.ndim == 3:
            result = result.argmax(-1)
        for i in range(result.shape[0]):
            text = "".join(id_to_word[int(t)] + " " for t in result[i])        # decode_caption: every word followed by ' '
            predictions.append({"p": text.split(separator, 1)[0], "r": captions[i][0].lower()})
    if cache_path:
        with open(cache_path, "wb") as f:
            pickle.dump(predictions, f, protocol=2)
    return predictions


def read_region_descriptions(data_file, image_ids=None):
    """Visual Genome region_descriptions.json -> {image_id: [(roi [y1,x1,y2,x2] px, phrase), ...]}
    (text_generation_model.py:57-80)."""
    with open(data_file, "r", encoding="utf-8") as doc:
        data = json.load(doc)
    want = None if image_ids is None else set(image_ids)
    out = {}
    for item in data:
        if want is not None and item["id"] not in want:
            continue
        out[item["id"]] = [([d["y"], d["x"], d["y"] + d["height"], d["x"] + d["width"]], d["phrase"])
                           for d in item["regions"]]
    return out


class RegionCaptionDataset(object):
    """The part of the reference's VisualGenomeDataset the caption models read (text_generation_model.py:50-128):
    per image the RoIs in pixels and their captions, encoded and framed to PADDING_SIZE."""

    def __init__(self, words_to_ids, padding_size, tokenizer=tokenize):
        self.word_to_id, self.padding_size, self.tokenizer = words_to_ids, int(padding_size), tokenizer
        self.image_info, self._image_ids, self.rois = {}, [], None

    def add_image(self, image_id, regions, **info):
        self.image_info[image_id] = dict(info, id=image_id, rois=[r for r, _ in regions], captions=[[p] for _, p in regions])
        self._image_ids.append(image_id)

    def load_visual_genome(self, data_file, image_ids=None):
        for image_id, regions in read_region_descriptions(data_file, image_ids).items():
            self.add_image(image_id, regions)

    @property
    def image_ids(self):
        return list(self._image_ids)

    def add_rois(self, rois):
        self.rois = rois

    def add_sequences(self, sequences):
        """v2 training sequences (load_sequences) -- text_generation_model_v2.py:95-96."""
        self.sequences = sequences

    def encode_region_caption(self, caption):
        return encode_caption(caption, self.word_to_id, self.tokenizer)

    def load_captions_and_rois(self, image_id):
        """-> rois [n, 4] (pixels), captions [n, P] float32; regions whose caption encodes to nothing are skipped."""
        info = self.image_info[image_id]
        rois, caps = [], []
        for roi, caption in zip(info["rois"], info["captions"]):
            ids = self.encode_region_caption(caption[0])
            if ids.size:
                rois.append(roi)
                caps.append(frame_caption(ids, self.padding_size))
        caps = np.stack(caps) if caps else np.zeros((0, self.padding_size), np.float32)
        return np.array(rois), caps

    def load_original_captions_and_rois(self, image_id):
        info = self.image_info[image_id]
        return np.array(info["rois"]), info["captions"]


def create_roi_info(dataset):
    """[(image_id, roi index within the image, framed caption)] over the dataset (text_generation_model.py:320-327)."""
    out = []
    for image_id in dataset.image_ids:
        _, captions = dataset.load_captions_and_rois(image_id)
        out.extend((image_id, i, captions[i]) for i in range(captions.shape[0]))
    return out


def targets_from_captions(captions, vocabulary_size=None):
    """The generator's targets: the caption shifted left by one with a trailing 0 (text_generation_model.py:352-357).
    Integer ids [..., P] by default (what the fused training step consumes); one-hot float64 [..., P, V] as the
    reference materialises them when `vocabulary_size` is given."""
    cap = np.asarray(captions)
    ids = np.concatenate([cap[..., 1:], np.zeros(cap.shape[:-1] + (1,), cap.dtype)], -1).astype(np.int32)
    if vocabulary_size is None:
        return ids
    onehot = np.zeros(ids.shape + (int(vocabulary_size),), np.float64)
    np.put_along_axis(onehot, ids[..., None].astype(np.int64), 1.0, -1)
    return onehot


def data_generator(dataset, features_fn, config, batch_size, shuffle=False, one_hot=False, shuffle_fn=None):
    """Endless v1 training batches (text_generation_model.py:330-372): ``dataset.rois`` is the create_roi_info list,
    ``features_fn(image_id) -> [n_rois, ...]`` the per-image RoI features (cached for consecutive RoIs of one image).
    Yields ([features [B, ...], input_words [B, P]], targets) with integer targets unless `one_hot`."""
    shuffle_fn = np.random.shuffle if shuffle_fn is None else shuffle_fn
    order = np.arange(len(dataset.rois))
    b, pos, prev_image, prev_features = 0, -1, None, None
    feats = words = None
    while True:
        pos = (pos + 1) % len(order)
        if shuffle and pos == 0:
            shuffle_fn(order)
        image_id, roi_index, cap = dataset.rois[order[pos]]
        if image_id != prev_image:
            prev_features, prev_image = features_fn(image_id), image_id
        f = np.asarray(prev_features[roi_index])
        if b == 0:
            feats = np.zeros((batch_size,) + f.shape, f.dtype)
            words = np.zeros((batch_size,) + np.shape(cap), np.asarray(cap).dtype)
        feats[b], words[b] = f, cap
        b += 1
        if b >= batch_size:
            yield [feats, words], targets_from_captions(words, config.VOCABULARY_SIZE if one_hot else None)
            b = 0


def load_sequences(dataset):
    """v2 training sequences (text_generation_model_v2.py:128-137) from integer captions: for caption c of RoI i,
    (image_id, i, [0], c[0]) and then (image_id, i, c[:j], c[j]) for j = 1 .. len(c) - 1.  `dataset` supplies
    ``load_captions_and_rois(image_id) -> (rois, captions)`` with captions as per-RoI id sequences."""
    out = []
    for image_id in dataset.image_ids:
        _, captions = dataset.load_captions_and_rois(image_id)
        for i, cap in enumerate(captions):
            cap = [int(c) for c in cap]
            out.append((image_id, i, [0], cap[0]))
            out.extend((image_id, i, cap[:j], cap[j]) for j in range(1, len(cap)))
    return out


def pad_sequences(sequences, maxlen, dtype="int32", value=0):
    """keras.preprocessing.sequence.pad_sequences with its defaults, as the v2 scripts call it
    (text_generation_model_v2.py:183, evaluate_models/test_score_dense_captions.py:219): padding='pre',
    truncating='pre' -- shorter sequences get ``value`` in FRONT, longer ones keep their LAST ``maxlen`` entries."""
    out = np.full((len(sequences), int(maxlen)), value, dtype=dtype)
    for i, s in enumerate(sequences):
        s = list(s)[-int(maxlen):] if maxlen > 0 else []
        if s:
            out[i, -len(s):] = s
    return out


def sequence_generator(dataset, features_fn, config, batch_size, shuffle=False, one_hot=True, shuffle_fn=None):
    """Batches of the v2 (next-word) model, the reference's second ``data_generator``
    (text_generation_model_v2.py:169-205): endless, one training sequence (image_id, roi index, previous words, next word)
    of ``dataset.sequences`` after the other, yields ``([features [B, p, p, C], previous words [B, PADDING_SIZE] int32,
    pre-padded], next word)`` where the next word is the reference's one-hot float64 ``[B, VOCABULARY_SIZE]`` row
    (``one_hot=True``) or the class id ``[B]`` int32 (InjectModelV2.train_on_batch takes either).  ``features_fn(image_id)``
    returns the RoI features of ALL RoIs of an image and is called once per run of sequences of the same image, as the
    reference caches ``generate_features``.  ``shuffle`` re-shuffles the sequence order at every wrap-around
    (``shuffle_fn`` defaults to ``np.random.shuffle``, the reference's)."""
    order = np.arange(len(dataset.sequences))
    shuffle_fn = shuffle_fn or np.random.shuffle
    b, idx, prev_image, prev_feats = 0, -1, None, None
    V, P = int(config.VOCABULARY_SIZE), int(config.PADDING_SIZE)
    while True:
        idx = (idx + 1) % len(order)
        if shuffle and idx == 0:
            shuffle_fn(order)
        image_id, roi_id, prev_words, next_word = dataset.sequences[order[idx]][:4]
        if prev_image != image_id:
            prev_feats = features_fn(image_id)
        prev_image = image_id
        roi_features = prev_feats[roi_id]
        if b == 0:
            batch_f = np.zeros((batch_size,) + roi_features.shape, dtype=roi_features.dtype)
            batch_w = np.zeros((batch_size, P), dtype=np.int32)
            batch_y = np.zeros((batch_size, V)) if one_hot else np.zeros((batch_size,), np.int32)
        batch_f[b] = roi_features
        batch_w[b] = pad_sequences([prev_words], P)[0]
        if one_hot:
            batch_y[b, int(next_word)] = 1
        else:
            batch_y[b] = int(next_word)
        b += 1
        if b >= batch_size:
            yield [batch_f, batch_w], batch_y
            b = 0


# ---------------------------------------------------------------------------------------------------
# Keras HDF5 weight files
# ---------------------------------------------------------------------------------------------------

def _text(v):
    return v.decode("utf8") if isinstance(v, bytes) else str(v)


def weights_from_h5_group(group):
    """Walk a Keras `save_weights` layout -- ``group.attrs['layer_names']``, then per layer
    ``group[layer].attrs['weight_names']`` naming the datasets -- into {'layer/weight': float32 array} with the
    ':0' suffix dropped.  Works on h5py groups and on any mapping with the same shape (tests use a stand-in).
    A full-model file keeps the same layout under 'model_weights'."""
    if "layer_names" not in group.attrs and "model_weights" in group:
        group = group["model_weights"]
    out = {}
    for layer in group.attrs["layer_names"]:
        g = group[_text(layer)]
        for name in g.attrs["weight_names"]:
            name = _text(name)
            key = name[:-2] if name.endswith(":0") else name
            out[key] = np.asarray(g[name], dtype=np.float32)
    return out


def load_keras_h5_weights(path):
    """{'layer/weight': array} from a Keras .h5 weight file (e.g. mask_rcnn_coco.h5 for the RoI head, or a
    text-model checkpoint).  Feed the result to ``model.set_weights(d, by_name=True)`` /
    ``RoiCaptionModel.load_weights``-style name matching."""
    try:
        import h5py
    except ImportError as e:                                  # pragma: no cover - h5py is absent in this image
        raise ImportError("reading Keras HDF5 weight files needs h5py, which is not installed here; convert the "
                          "file offline to .npz keyed by Keras weight names (RoiCaptionModel.load_weights reads that)") from e
    with h5py.File(path, "r") as f:
        return weights_from_h5_group(f)


def write_weight_file(path, weights):
    """{'layer/weight': array} -> an .npz archive keyed by Keras weight name, written EXACTLY at ``path`` whatever its
    extension: the reference's scripts name their checkpoints ``*.h5`` (ModelCheckpoint(model_filepath, ...),
    text_generation_model.py:461) and read them back under the same name (:484)."""
    with open(path, "wb") as f:
        np.savez(f, **{k.replace("/", "__"): np.asarray(v) for k, v in weights.items()})


def read_weight_file(path):
    """{'layer/weight': array} from a weight file, the format decided by the file's first bytes and not by its name:
    an .npz archive as written by ``write_weight_file`` / ``save_weights`` (zip magic), or Keras HDF5 (needs h5py)."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        return load_keras_h5_weights(path)
    if magic[:2] == b"PK":
        with np.load(path) as z:
            return {k.replace("__", "/"): z[k] for k in z.files}
    raise ValueError("%s is neither an .npz archive nor an HDF5 file" % path)


def convert_keras_h5_to_npz(h5_path, npz_path):
    """Offline converter: Keras .h5 weights -> the .npz (keyed by Keras weight name) that load_weights reads."""
    np.savez(npz_path, **{k.replace("/", "__"): v for k, v in load_keras_h5_weights(h5_path).items()})
