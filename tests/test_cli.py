"""Host-side tests of the CLI mirror (ladi_vton_b200/inference.py <- /root/reference/src/inference.py): flag-for-flag parity with the
reference's parser (golden read from the reference source by tests/golden/make_cli_golden.py), dataroot checks, prompt template,
output tree, stand-in tokenizer, rank sharding rule, and the no-CPU-path guarantee.  No GPU needed."""
import json
import os

import pytest
import torch

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_flags.json")))
BASE = ["--output_dir", "o", "--test_order", "paired", "--dataset", "vitonhd"]


def _actions():
    from ladi_vton_b200.inference import build_parser
    p = build_parser()
    return {a.option_strings[0]: a for a in p._actions if a.option_strings and a.option_strings[0] != "-h"}, p


def test_flags_match_reference_parser():
    acts, parser = _actions()
    engine = {a.option_strings[0] for g in parser._action_groups if g.title == "engine" for a in g._group_actions}
    assert engine == {"--checkpoint_dir", "--vision_model_name_or_path", "--reference_src", "--random_init", "--synthetic_samples"}
    ref_names = [f["flags"][0] for f in GOLD["flags"]]
    assert len(ref_names) == 18
    assert sorted(set(acts) - engine) == sorted(ref_names)  # nothing missing, nothing extra outside the engine group
    for f in GOLD["flags"]:
        a = acts[f["flags"][0]]
        assert a.option_strings == f["flags"]
        if f.get("action") == "store_true":
            assert a.nargs == 0 and a.const is True and a.default is False
        else:
            assert a.type.__name__ == f["type"], f
            assert a.default == f.get("default"), f
        assert bool(a.required) == bool(f.get("required", False)), f
        assert (list(a.choices) if a.choices else None) == f.get("choices"), f


def test_defaults_and_required():
    from ladi_vton_b200.inference import parse_args
    a = parse_args(BASE)
    assert (a.seed, a.batch_size, a.num_workers, a.num_vstar, a.num_inference_steps, a.guidance_scale) == (1234, 8, 8, 16, 50, 7.5)
    assert a.category == "all" and a.mixed_precision is None and not a.use_png and not a.compute_metrics
    for drop in ("--output_dir", "--test_order", "--dataset"):
        argv = list(BASE)
        i = argv.index(drop)
        del argv[i:i + 2]
        with pytest.raises(SystemExit):
            parse_args(argv)
    with pytest.raises(SystemExit):
        parse_args(BASE + ["--category", "shoes"])


def test_dataroot_checks():
    """src/inference.py:104-108."""
    from ladi_vton_b200.inference import check_args, parse_args
    with pytest.raises(ValueError, match="VitonHD dataroot must be provided"):
        check_args(parse_args(BASE))
    with pytest.raises(ValueError, match="DressCode dataroot must be provided"):
        check_args(parse_args(["--output_dir", "o", "--test_order", "unpaired", "--dataset", "dresscode"]))
    check_args(parse_args(BASE + ["--vitonhd_dataroot", "/data"]))
    check_args(parse_args(BASE + ["--synthetic_samples", "4"]))


def test_prompt_template_and_tokenizer():
    from ladi_vton_b200.inference import CATEGORY_TEXT, StandInTokenizer, prompts_for
    assert CATEGORY_TEXT == GOLD["category_text"]
    p = prompts_for(["dresses", "upper_body", "lower_body"], 16)
    assert p[0] == "a photo of a model wearing a dress " + " $ " * 16
    assert p[1].startswith("a photo of a model wearing an upper body garment  $ ") and p[2].count("$") == 16
    tok = StandInTokenizer()
    ids = tok(p, max_length=tok.model_max_length, padding="max_length", truncation=True, return_tensors="pt").input_ids
    assert ids.shape == (3, 77) and ids.dtype == torch.long
    assert (ids[:, 0] == 49406).all() and (ids[:, -1] == 49407).all()
    assert ((ids == 259).sum(1) == 16).all()
    first = [int((r == 259).nonzero()[0]) for r in ids]
    for r, f in zip(ids, first):  # the 16 placeholders are contiguous (encode_text_word_embedding writes num_vstar rows from the first)
        assert (r[f:f + 16] == 259).all()
    assert torch.equal(tok("a $ b").input_ids, tok(["a $ b"]).input_ids)
    long = tok(" ".join(["w"] * 200)).input_ids
    assert long.shape == (1, 77) and long[0, -1] == 49407


def test_output_tree(tmp_path):
    """src/inference.py:222-224,314-324: output_dir/<test_order>/<category>/<im_name>, jpg (quality 95) or png."""
    from PIL import Image
    from ladi_vton_b200.inference import save_images
    ims = [Image.new("RGB", (8, 6), (i * 40, 0, 0)) for i in range(3)]
    cats, names = ["upper_body", "dresses", "upper_body"], ["000_00.jpg", "001_00.jpg", "002_00.jpg"]
    save_dir = os.path.join(tmp_path, "unpaired")
    paths = save_images(ims, cats, names, save_dir, use_png=False)
    assert sorted(os.listdir(save_dir)) == ["dresses", "upper_body"]
    assert sorted(os.listdir(os.path.join(save_dir, "upper_body"))) == ["000_00.jpg", "002_00.jpg"]
    assert Image.open(paths[1]).format == "JPEG"
    paths = save_images(ims, cats, names, save_dir, use_png=True)
    assert paths[0].endswith("000_00.png") and Image.open(paths[0]).format == "PNG"
    assert Image.open(paths[2]).getpixel((0, 0)) == (80, 0, 0)  # png is lossless


def test_synthetic_dataset_has_reference_batch_keys():
    from ladi_vton_b200.inference import OUTPUTLIST, SyntheticTryOnDataset
    ds = SyntheticTryOnDataset(3, size=(64, 48), categories=["dresses", "upper_body"])
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)))
    assert sorted(batch) == sorted(OUTPUTLIST)
    assert batch["image"].shape == batch["cloth"].shape == batch["im_mask"].shape == (2, 3, 64, 48)
    assert batch["pose_map"].shape == (2, 18, 64, 48) and batch["inpaint_mask"].shape == (2, 1, 64, 48)
    assert set(batch["inpaint_mask"].unique().tolist()) <= {0.0, 1.0}
    assert batch["category"] == ["dresses", "upper_body"] and batch["im_name"][1] == "00001_00.jpg"
    assert float(batch["image"].abs().max()) <= 1 and float(batch["cloth"].abs().max()) <= 1
    assert float((batch["im_mask"] * batch["inpaint_mask"]).abs().max()) == 0.0  # the agnostic image is blank under the inpaint mask


def test_missing_dataset_classes_error():
    from ladi_vton_b200.inference import build_dataset, parse_args
    a = parse_args(BASE + ["--vitonhd_dataroot", "/nonexistent"])
    with pytest.raises(ImportError, match="--reference_src"):
        build_dataset(a, ["upper_body"])


def test_missing_weights_error(tmp_path):
    from ladi_vton_b200.inference import _load_folder_state_dict
    with pytest.raises(FileNotFoundError, match="no network"):
        _load_folder_state_dict(str(tmp_path))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_cli_has_no_cpu_path(tmp_path):
    from ladi_vton_b200.inference import main
    with pytest.raises(RuntimeError, match="no CPU path"):
        main(["--output_dir", str(tmp_path), "--test_order", "paired", "--dataset", "vitonhd", "--synthetic_samples", "2", "--random_init"])
    assert os.listdir(tmp_path) == []


def test_rank_sharding_covers_every_batch_once():
    from ladi_vton_b200.inference import batches_for_rank
    for n, world in ((7, 2), (8, 8), (3, 4), (0, 2), (10, 1)):
        parts = [batches_for_rank(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert batches_for_rank(7, 1, 2) == [1, 3, 5]
