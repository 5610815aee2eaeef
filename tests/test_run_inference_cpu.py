"""CPU: host-side helpers of the run_inference driver (file discovery and naming follow dataset/custom_dataset.py:16-19,29)."""
import os

from smap_b200 import run_inference as R


def test_list_images_and_names(tmp_path):
    d = tmp_path / "data"
    (d / "a" / "b").mkdir(parents=True)
    for rel in ("z.jpg", "a/x.png", "a/b/y.jpeg", "a/skip.txt", "a/b/w.JPG"):
        (d / rel).write_bytes(b"")
    got = R.list_images(str(d))
    assert got == sorted(got)
    rels = [R.image_name(p, str(d)) for p in got]
    assert rels == ["a/b/y.jpeg", "a/x.png", "z.jpg"]          # jpg / png / jpeg only, case-sensitive like glob
    assert R.image_name(str(d) + "/q.png\n", str(d)) == "q.png"


def test_cli_reports_missing_checkpoint(tmp_path, capsys):
    rc = R.main(["-p", os.path.join(tmp_path, "nope.pth"), "--dataset_path", str(tmp_path)])
    assert rc == 1
    assert "No such checkpoint of SMAP" in capsys.readouterr().out


def test_glob_order_option_keeps_the_reference_order(tmp_path):
    """dataset/custom_dataset.py:16-18 concatenates one glob per extension (jpg, png, jpeg) without sorting."""
    import glob

    d = tmp_path / "data"
    (d / "s").mkdir(parents=True)
    for rel in ("b.png", "a.jpg", "s/c.jpeg", "s/0.jpg"):
        (d / rel).write_bytes(b"")
    want = []
    for ext in ("jpg", "png", "jpeg"):
        want.extend(glob.glob(os.path.join(str(d), "**/*." + ext), recursive=True))
    assert R.list_images(str(d), glob_order=True) == want
    assert sorted(want) == R.list_images(str(d))
