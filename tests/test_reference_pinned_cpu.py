"""Host logic pinned to the REFERENCE'S OWN CODE.  The fixtures under tests/golden/ were produced by
scripts/gen_golden_agent.py, which imports /root/reference (unmodified) in the build container and runs
  * `VLNEvaluator.step` (streamvln/streamvln_agent.py:169-258) for 104 env steps against a recording model,
  * `StreamVLNForCausalLM.prepare_inputs_for_generation` (streamvln/model/stream_video_vln.py:409-471),
  * the geometry helpers of the Habitat evaluator (streamvln/streamvln_eval.py:126-185).
Checked here, on CPU: the oracle's restatements (oracle/glue.py `OracleAgent`, the tail rule of `ReferenceOracle`) AND the
product's host side (`streamvln_b200.agent.VLNEvaluator.prepare/finish`, `model.uncached_tail`, `synth` geometry)."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden"


def _rows(z, key):
    return [z[key][i, : z[key + "_len"][i]].tolist() for i in range(z[key].shape[0])]


@pytest.fixture(scope="module")
def agent_gold():
    z = np.load(GOLD / "ref_agent_104steps.npz")
    return dict(z=z, inputs=_rows(z, "inputs"), views=_rows(z, "views"), time_ids=_rows(z, "time_ids"),
                actions=_rows(z, "actions"), instruction=bytes(z["instruction"]).decode(), n_steps=int(z["n_steps"]),
                resets=z["resets"].tolist(), fresh=z["fresh"].tolist(), texts=[t.decode() for t in z["texts"]])


def _seq(t, eos):
    return [(t * 7 + j * 3) % 50 + 10 for j in range(4)] + [eos]      # what the recording model returned


def test_fixture_is_the_expected_episode(agent_gold):
    g = agent_gold
    assert g["n_steps"] == 104 and len(g["inputs"]) == 26
    assert [r[0] for r in g["resets"]] == [8, 16, 24]                 # window boundaries after env steps 31, 63, 95
    assert g["views"][8] == [0, 4, 8, 12, 16, 20, 24, 28, 32] and g["views"][24] == [0, 12, 24, 36, 48, 60, 72, 84, 96]
    assert g["fresh"] == [i % 8 == 0 for i in range(26)]
    assert all(m == 10000 for m in g["z"]["max_new_tokens"].tolist())            # streamvln_agent.py:244


def test_oracle_agent_is_pinned_to_the_reference_step(agent_gold):
    from oracle import glue
    from streamvln_b200.tokenizer import StubQwenTokenizer

    g = agent_gold
    tok = StubQwenTokenizer(152064)
    calls, resets = [], []

    class FakeOracle:
        step_logits = []

        def reset_for_env(self, idx):
            resets.append(len(calls))

        def generate(self, ids, feats, time_ids, idx, past, max_new, force_tokens=None):
            t = len(calls)
            calls.append(dict(inputs=list(ids), views=feats[:, 0, 0].to(torch.int64).tolist(), time_ids=list(time_ids[0]),
                              fresh=past is None, max_new=max_new))
            return torch.tensor([_seq(t, tok.im_end)]), ("kv", t), torch.zeros(1)

    oa = glue.OracleAgent(FakeOracle(), tok, lambda i: torch.full((1, 1), float(i)), 32, 4, 8)
    actions = []
    for s in range(g["n_steps"]):
        a = oa.step(0, g["instruction"], run_model=(s % 4 == 0))
        actions.append([-1] if a is None else a)
        oa.step_id += 1
    assert [c["inputs"] for c in calls] == g["inputs"]
    assert [c["views"] for c in calls] == g["views"]
    assert [c["time_ids"] for c in calls] == g["time_ids"]
    assert [c["fresh"] for c in calls] == g["fresh"]
    assert resets == [r[0] for r in g["resets"]]
    assert actions == g["actions"]
    assert all(c["max_new"] == 10000 for c in calls)


def test_product_agent_is_pinned_to_the_reference_step(agent_gold):
    """streamvln_b200.agent.VLNEvaluator (host side: buffering, window reset, history slice, prompt ids, parse) against
    the same recording-model protocol.  `frame_ids` - the product's extra kwarg - must name exactly the views the
    reference stacked into `images`."""
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.tokenizer import StubQwenTokenizer

    g = agent_gold
    tok = StubQwenTokenizer(152064)
    calls, resets = [], []

    class RecordingModel:
        device = torch.device("cpu")
        config = SimpleNamespace(mode="reference")

        def get_vision_tower(self):
            return SimpleNamespace(image_processor=SimpleNamespace(crop_size={"height": 384, "width": 384}))

        def reset_for_env(self, idx):
            resets.append(len(calls))

        def generate(self, **kw):
            t = len(calls)
            calls.append(dict(inputs=kw["inputs"][0].tolist(), views=kw["images"][0, :, 0, 0, 0].to(torch.int64).tolist(),
                              frame_ids=list(kw["frame_ids"]), time_ids=list(kw["time_ids"][0]), fresh=kw["past_key_values"] is None,
                              max_new=kw["max_new_tokens"], shapes=[tuple(kw[k].shape) for k in ("images", "depths", "poses", "intrinsics")],
                              flags=[kw["do_sample"], kw["num_beams"], kw["use_cache"], kw["return_dict_in_generate"]],
                              task_ids=kw["task_ids"], env_id=kw["env_id"]))
            return SimpleNamespace(sequences=torch.tensor([_seq(t, tok.im_end)]), past_key_values=("kv", t))

    K = np.array([[192.0, 0, 191.42857143, 0], [0, 192.0, 191.42857143, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, gpu_preprocess=False)
    ev = VLNEvaluator({"rgb_height": 1.25, "camera_intrinsic": K}, model=RecordingModel(), tokenizer=tok, args=args)
    actions, texts = [], []
    for s in range(g["n_steps"]):
        stamp = torch.full((3, 2, 2), float(s % 256))               # what the fixture's stub image processor produced
        a, _, txt = ev.step(0, None, g["instruction"], run_model=(s % 4 == 0), preprocessed=stamp if s % 4 == 0 else None)
        actions.append([-1] if a is None else a)
        texts.append("" if txt is None else txt)
        ev.step_id += 1
    assert [c["inputs"] for c in calls] == g["inputs"]
    assert [c["views"] for c in calls] == g["views"] == [c["frame_ids"] for c in calls]
    assert [c["time_ids"] for c in calls] == g["time_ids"]
    assert [c["fresh"] for c in calls] == g["fresh"]
    assert resets == [r[0] for r in g["resets"]]
    assert actions == g["actions"] and texts == g["texts"]
    assert all(c["max_new"] == 10000 and c["flags"] == [False, 1, True, True] and c["task_ids"] == [0] and c["env_id"] == 0 for c in calls)
    # window start: 8 history views + the current frame, geometry stacked alongside (streamvln_agent.py:219-236)
    assert calls[8]["shapes"][0][:2] == (1, 9) and calls[8]["shapes"][2] == (1, 9, 4, 4) and calls[8]["shapes"][3] == (1, 9, 4, 4)
    assert tuple(g["z"]["pose_shape"]) == calls[8]["shapes"][2]


def test_tail_rule_is_pinned_to_prepare_inputs_for_generation():
    from streamvln_b200.model import uncached_tail

    z = np.load(GOLD / "ref_tail_rule.npz")
    names = [n.decode() for n in z["names"]]
    for i, name in enumerate(names):
        total, past = int(z["n_embeds"][i]), int(z["past_len"][i])
        assert bool(z["position_ids_none"][i])                      # stream_video_vln.py:460: positions are absolute indices
        if z["uses_embeds"][i]:
            # product
            assert uncached_tail(total, past) == (int(z["first_row"][i]), int(z["n_rows"][i])), name
            # oracle (glue.ReferenceOracle.generate runs full[:, past_len:])
            assert total - past == int(z["n_rows"][i]) and past == int(z["first_row"][i])
        elif name == "decode":
            assert int(z["ids_out"][i]) == 100 + int(z["n_ids"][i]) - 1   # the last generated id is what is fed
    # the reference's own edge: ONE un-cached embed row falls through to the input_ids branch (`len(cache_position) > 1`
    # is false) and would feed an empty id tensor; no turn is that short (>= 200 rows), the product runs the row
    i = names.index("one_uncached")
    assert not z["uses_embeds"][i] and uncached_tail(300, 299) == (299, 1)
    with pytest.raises(ValueError):
        uncached_tail(300, 300)


def test_synth_geometry_is_pinned_to_the_habitat_evaluator():
    from streamvln_b200 import synth

    z = np.load(GOLD / "ref_geometry.npz")
    np.testing.assert_allclose(synth.intrinsic_matrix(640, 480, 79.0, 384), z["K384"].astype(np.float32), rtol=0, atol=0)
    np.testing.assert_array_equal(synth.AXIS_ALIGN, z["axis_align"])
    for i in range(z["xyz"].shape[0]):
        x, y, zz = z["xyz"][i]
        # the eval loop passes (x, -y, z) (streamvln_eval.py:261-262); pose_matrix takes the habitat y and flips it itself
        got = synth.pose_matrix(x, -y, float(z["yaw"][i]), cam_h=zz)
        np.testing.assert_array_equal(got, z["pose"][i].astype(np.float32))


def test_depth_preprocessing_is_pinned_to_the_habitat_evaluator():
    from streamvln_b200 import synth

    z = np.load(GOLD / "ref_geometry.npz")
    got = synth.preprocess_depth(z["depth01"])
    assert got.dtype == np.float32 and got.shape == (384, 384)
    np.testing.assert_array_equal(got, z["depth384"].astype(np.float32))
