"""Host-side data formats (image_captioning_b200/data.py) against the reference's own code where it runs here
(tests/golden/reference_numpy.npz: load_corpus, encode_caption's token filtering, data_generator, load_sequences)
and against the layouts the reference reads (region_descriptions.json, the vocabulary pickles, Keras HDF5 groups).
CPU only."""
import json
import os
import pickle

import numpy as np
import pytest

from image_captioning_b200 import data

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_numpy.npz"))
TOKENS = ["a", "man", "dog", "red", "on", "the", "grass", "frisbee", "with", "'s", "two", "playing"]


def _vocab():
    words = [str(w) for w in G["corpus_words"]]
    return {w: i for i, w in enumerate(words)}, dict(enumerate(words))


def test_load_corpus_equals_the_reference():
    emb = {w: G["corpus_matrix"][i + 3] for i, w in enumerate(TOKENS)}
    np.random.seed(11)
    w2i, i2w, mat = data.load_corpus(TOKENS, emb, 6)
    assert np.array_equal(mat, G["corpus_matrix"])
    assert [i2w[i] for i in range(len(i2w))] == [str(w) for w in G["corpus_words"]]
    assert w2i["<unk>"] == 0 and w2i["<start>"] == 1 and w2i["<end>"] == 2 and w2i["a"] == 3
    assert not mat[0].any() and (np.abs(mat[1:3]) <= 0.5).all()


def test_encode_caption_drops_unknown_words_like_the_reference():
    w2i, i2w = _vocab()
    for text, want in zip(G["captions_text"], G["captions_encoded"]):
        got = data.encode_caption(str(text), w2i)
        assert ",".join(str(int(v)) for v in got) == str(want)
    assert data.tokenize("A man's dog, isn't it?") == ["a", "man", "'s", "dog", ",", "is", "n't", "it", "?"]
    assert data.decode_caption([1, 3, 4, 2, 0], i2w) == "<start> a man <end> <unk>"
    assert data.decode_caption([3, 4, 2, 0], i2w, stop="<end>") == "a man"


def test_frame_caption_start_end_cut_and_post_padding():
    assert data.frame_caption([5, 6], 6).tolist() == [1, 5, 6, 2, 0, 0]
    assert data.frame_caption([5, 6, 7, 8, 9, 10], 6).tolist() == [1, 5, 6, 7, 8, 2]      # cut to P - 2 words
    assert data.frame_caption([5, 6, 7, 8], 6).tolist() == [1, 5, 6, 7, 8, 2]
    assert data.frame_caption([5], 6).dtype == np.float32


def test_region_descriptions_dataset_and_generator(tmp_path):
    w2i, _ = _vocab()
    doc = [{"id": 7, "regions": [{"x": 10, "y": 20, "width": 30, "height": 40, "phrase": "a man on the grass"},
                                 {"x": 1, "y": 2, "width": 3, "height": 4, "phrase": "zebra"},
                                 {"x": 5, "y": 6, "width": 7, "height": 8, "phrase": "Red dog."}]},
           {"id": 9, "regions": [{"x": 0, "y": 0, "width": 9, "height": 9, "phrase": "two playing with a frisbee"}]},
           {"id": 11, "regions": []}]
    path = tmp_path / "region_descriptions.json"
    path.write_text(json.dumps(doc))
    ds = data.RegionCaptionDataset(w2i, 6)
    ds.load_visual_genome(str(path), image_ids=[7, 9])
    assert ds.image_ids == [7, 9]
    rois, caps = ds.load_captions_and_rois(7)
    assert rois.tolist() == [[20, 10, 60, 40], [6, 5, 14, 12]]            # the out-of-vocabulary region is skipped
    assert caps.tolist() == [[1, 3, 4, 7, 8, 2], [1, 6, 5, 2, 0, 0]] and caps.dtype == np.float32
    rois_o, caps_o = ds.load_original_captions_and_rois(7)
    assert rois_o.shape == (3, 4) and caps_o[1] == ["zebra"]
    ds.add_rois(data.create_roi_info(ds))
    assert [(a, b) for a, b, _ in ds.rois] == [(7, 0), (7, 1), (9, 0)]
    calls = []

    def feats(image_id):
        calls.append(image_id)
        return np.full((2, 7, 7, 4), image_id, np.float32)

    class Cfg:
        VOCABULARY_SIZE = len(w2i)
    gen = data.data_generator(ds, feats, Cfg, 2)
    (f, words), tgt = next(gen)
    assert f.shape == (2, 7, 7, 4) and words.tolist() == caps.tolist() and calls == [7]     # one extraction per image
    assert tgt.tolist() == [[3, 4, 7, 8, 2, 0], [6, 5, 2, 0, 0, 0]] and tgt.dtype == np.int32


def test_data_generator_equals_the_reference():
    caps = {7: G["gen_caps_7"], 9: G["gen_caps_9"]}
    feats = {7: G["gen_feats_7"], 9: G["gen_feats_9"]}

    class DS:
        image_ids = [7, 9]
        rois = [(k, i, caps[k][i]) for k in (7, 9) for i in range(caps[k].shape[0])]

        def load_captions_and_rois(self, image_id):
            return None, caps[image_id]

    class Cfg:
        VOCABULARY_SIZE = G["gen_targets"].shape[-1]
    gen = data.data_generator(DS(), lambda i: feats[i], Cfg, 2, one_hot=True)
    ids = data.data_generator(DS(), lambda i: feats[i], Cfg, 2)
    for b in range(3):
        (f, w), t = next(gen)
        assert np.array_equal(f, G["gen_features"][b]) and np.array_equal(w, G["gen_words"][b])
        assert np.array_equal(t, G["gen_targets"][b]) and t.dtype == G["gen_targets"].dtype
        assert np.array_equal(next(ids)[1], t.argmax(-1))
    seqs = data.load_sequences(DS())
    got = ["%d|%d|%s|%d" % (a, b, ",".join(str(int(v)) for v in c), d) for a, b, c, d in seqs]
    assert got == [str(v) for v in G["v2_sequences"]]


def test_vocabulary_pickles_round_trip(tmp_path):
    w2i, i2w = _vocab()
    paths = [str(tmp_path / n) for n in ("id_to_word.pickle", "word_to_id.pickle", "embedding_matrix.pickle")]
    for p, obj in zip(paths, (i2w, w2i, G["corpus_matrix"])):
        with open(p, "wb") as f:
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
    a, b, m = data.load_vocabulary(*paths)
    assert a == w2i and b == i2w and np.array_equal(m, G["corpus_matrix"])


class _Node(dict):
    def __init__(self, items=(), **attrs):
        super().__init__(items)
        self.attrs = attrs


def test_keras_h5_group_walker_on_a_stand_in():
    k1 = np.arange(12, dtype=np.float64).reshape(3, 4)
    layer = _Node({"mrcnn_class_conv1/kernel:0": k1, "mrcnn_class_conv1/bias:0": np.ones(4)},
                  weight_names=[b"mrcnn_class_conv1/kernel:0", b"mrcnn_class_conv1/bias:0"])
    lstm = _Node({"imgcap_lstm1/kernel:0": np.zeros((2, 8))}, weight_names=[b"imgcap_lstm1/kernel:0"])
    root = _Node({"mrcnn_class_conv1": layer, "imgcap_lstm1": lstm, "input_1": _Node(weight_names=[])},
                 layer_names=[b"input_1", b"mrcnn_class_conv1", b"imgcap_lstm1"])
    for top in (root, _Node({"model_weights": root})):
        w = data.weights_from_h5_group(top)
        assert sorted(w) == ["imgcap_lstm1/kernel", "mrcnn_class_conv1/bias", "mrcnn_class_conv1/kernel"]
        assert w["mrcnn_class_conv1/kernel"].dtype == np.float32 and np.array_equal(w["mrcnn_class_conv1/kernel"], k1)
    try:
        import h5py  # noqa: F401
    except ImportError:
        try:
            data.load_keras_h5_weights("missing.h5")
        except ImportError as e:
            assert "h5py" in str(e)
        else:
            raise AssertionError("expected ImportError without h5py")


def test_weight_files_are_told_apart_by_content_not_by_name(tmp_path):
    """A checkpoint the reference's script names *.h5 (ModelCheckpoint(model_filepath), text_generation_model.py:461) is
    written as an .npz archive at exactly that path and read back under the same name."""
    from image_captioning_b200 import data
    w = {"imgcap_lstm1/kernel": np.arange(12, dtype=np.float32).reshape(3, 4), "mrcnn_class_bn1/gamma": np.ones(5, np.float32)}
    path = str(tmp_path / "weights-01.h5")
    data.write_weight_file(path, w)
    import os
    assert os.path.exists(path) and not os.path.exists(path + ".npz")
    back = data.read_weight_file(path)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    junk = str(tmp_path / "junk.h5")
    with open(junk, "wb") as f:
        f.write(b"not a weight file")
    with pytest.raises(ValueError):
        data.read_weight_file(junk)
    hdf = str(tmp_path / "real.h5")
    with open(hdf, "wb") as f:
        f.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(ImportError):                        # HDF5 by content: needs h5py, which this image lacks
        data.read_weight_file(hdf)


def test_v2_sequence_generator_equals_the_reference():
    """data.sequence_generator against the v2 script's own data_generator (text_generation_model_v2.py:169-205), run on
    the same fake dataset by tests/golden/gen_golden_reference_numpy.py: 18 sequences in batches of 4 (one wrap-around),
    PADDING_SIZE 4 < longest prefix (pre-truncation), features fetched once per run of sequences of an image."""
    caps = {7: G["gen_caps_7"], 9: G["gen_caps_9"]}
    feats = {7: G["gen_feats_7"], 9: G["gen_feats_9"]}

    class DS:
        image_ids = [7, 9]

        def load_captions_and_rois(self, image_id):
            return None, caps[image_id]
    ds = DS()
    ds.sequences = data.load_sequences(ds)
    assert len(ds.sequences) == 18

    class Cfg:
        VOCABULARY_SIZE = G["v2gen_next"].shape[-1]
        PADDING_SIZE = G["v2gen_words"].shape[-1]
    calls = []

    def fetch(image_id):
        calls.append(image_id)
        return feats[image_id]
    gen = data.sequence_generator(ds, fetch, Cfg, 4)
    ids = data.sequence_generator(ds, lambda i: feats[i], Cfg, 4, one_hot=False)
    for b in range(5):
        (f, w), y = next(gen)
        assert np.array_equal(f, G["v2gen_features"][b]) and f.dtype == G["v2gen_features"].dtype
        assert np.array_equal(w, G["v2gen_words"][b]) and w.dtype == G["v2gen_words"].dtype
        assert np.array_equal(y, G["v2gen_next"][b]) and y.dtype == G["v2gen_next"].dtype
        yi = next(ids)[1]
        assert yi.dtype == np.int32 and np.array_equal(yi, y.argmax(-1))
    assert calls == [7, 9, 7]                                # 12 sequences of image 7, 6 of image 9, then the wrap-around
    # keras pad_sequences defaults: pre-padding, pre-truncation
    assert data.pad_sequences([[0], [3, 7], [1, 2, 3, 4, 5, 6]], 4).tolist() == [[0, 0, 0, 0], [0, 0, 3, 7], [3, 4, 5, 6]]
    # shuffling happens at every wrap-around, with the reference's np.random.shuffle by default
    order = []
    sh = data.sequence_generator(ds, lambda i: feats[i], Cfg, 18, shuffle=True, shuffle_fn=lambda a: (order.append(1), a.__setitem__(slice(None), a[::-1].copy())))
    (f1, w1), y1 = next(sh)
    assert len(order) == 1 and np.array_equal(y1[0], G["v2gen_next"].reshape(-1, y1.shape[-1])[17])


def test_generate_predictions_is_the_reference_evaluation_loop(tmp_path):
    """data.generate_predictions (evaluate_models/eval_text_generation_model.py:129-153) with a stand-in model: captions
    decoded word by word, cut at ' <end>', paired with the lower-cased ground truth; pickle cache read back."""
    id_to_word = {0: "<pad>", 1: "<start>", 2: "<end>", 3: "a", 4: "red", 5: "dog"}

    class DS:
        image_ids = [7, 9]

        def load_original_captions_and_rois(self, image_id):
            return None, {7: [["A Red Dog"], ["Dog"]], 9: [["a DOG"]]}[image_id]

    class Model:
        calls = []

        def predict(self, features, batch_size=None):
            self.calls.append((features.shape[0], batch_size))
            ids = {2: [[3, 4, 5, 2, 0], [5, 2, 2, 0, 0]], 1: [[3, 5, 4, 3, 5]]}[features.shape[0]]
            ids = np.array(ids)
            return np.eye(6)[ids] if features.shape[0] == 1 else ids          # distributions [N,P,V] or ids [N,P]

    class Cfg:
        BATCH_SIZE = 4
    feats = {7: np.zeros((2, 7, 7, 4), np.float32), 9: np.zeros((1, 7, 7, 4), np.float32)}
    cache = str(tmp_path / "m_predictions.pickle")
    m = Model()
    got = data.generate_predictions(m, DS(), lambda i: feats[i], Cfg, id_to_word, cache_path=cache)
    assert got == [{"p": "a red dog", "r": "a red dog"}, {"p": "dog", "r": "dog"}, {"p": "a dog red a dog ", "r": "a dog"}]
    assert m.calls == [(2, 4), (1, 4)]
    again = data.generate_predictions(None, None, None, None, None, cache_path=cache)      # served from the cache
    assert again == got
