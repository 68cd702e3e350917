"""Tracking / registry (reference MLflow usage: P1/03:361-373,438; P2/01:221-299)."""
import math
import os

import pytest

from b200ddl import tracking


@pytest.fixture(autouse=True)
def store(tmp_path):
    tracking.set_tracking_uri(str(tmp_path / "mlruns"))
    tracking.set_experiment("/Users/me/distributed_dl_workshop")
    yield
    while tracking.active_run() is not None:
        tracking.end_run()


def test_run_lifecycle_and_logging():
    with tracking.start_run(run_name="horovod_driver") as run:
        rid = tracking.active_run().info.run_id
        tracking.log_param("epochs", 3)
        tracking.log_metric("val_loss", 1.5)
        tracking.log_metric("val_loss", 0.7)
        tracking.log_metrics({"val_accuracy": 0.9})
        tracking.log_dict({"img_height": 224, "img_width": 224}, "img_params_dict.json")
    tracking.end_run()  # no-op after a with-block, like the reference (Q9)
    r = tracking.get_run(rid)
    assert r.info.status == "FINISHED"
    assert r.data.params == {"epochs": "3"}
    assert r.data.metrics["val_loss"] == 0.7 and tracking.metric_history(rid, "val_loss") == [1.5, 0.7]
    assert os.path.exists(tracking.resolve_uri(f"runs:/{rid}/img_params_dict.json"))


def test_nested_runs_search_and_missing_metric_order():
    with tracking.start_run(run_name="hyperopt_tuning") as parent:
        pid = parent.info.run_id
        for i, acc in enumerate([0.3, 0.9, None]):
            with tracking.start_run(run_name=f"t{i}", nested=True):
                tracking.log_param("i", i)
                if acc is not None:
                    tracking.log_metric("accuracy", acc)
    df = tracking.search_runs(filter_string=f'tags.mlflow.parentRunId = "{pid}"', order_by=["metrics.accuracy DESC"])
    assert len(df) == 3
    assert df.iloc[0]["metrics.accuracy"] == 0.9
    assert math.isnan(df.iloc[2]["metrics.accuracy"])          # missing metric sorts last (Q3)
    # ordering by a metric nobody logged must not raise
    df2 = tracking.search_runs(filter_string=f'tags.mlflow.parentRunId = "{pid}"', order_by=["metrics.nope DESC"])
    assert len(df2) == 3


def test_resume_run_by_id_from_worker():
    with tracking.start_run(run_name="horovod_distributed") as run:
        rid = run.info.run_id
    # rank 0 re-opens the driver's run (reference P1/03:363)
    with tracking.start_run(run_id=rid):
        tracking.log_metric("val_accuracy", 0.5)
    assert tracking.get_run(rid).data.metrics["val_accuracy"] == 0.5


def test_registry_stage_transitions(tmp_path):
    with tracking.start_run() as run:
        tracking.log_dict({"a": 1}, "model/MLmodel.json")
        rid = run.info.run_id
    mv = tracking.register_model(f"runs:/{rid}/model", "me_flower_classifier")
    assert mv.version == 1 and mv.current_stage == "None"
    client = tracking.MlflowClient()
    client.transition_model_version_stage("me_flower_classifier", mv.version, stage="Production")
    p = tracking.resolve_uri("models:/me_flower_classifier/production")
    assert p.endswith(os.path.join(rid, "artifacts", "model"))
    mv2 = tracking.register_model(f"runs:/{rid}/model", "me_flower_classifier")
    assert mv2.version == 2
    assert tracking.resolve_uri("models:/me_flower_classifier/2") == p
    with pytest.raises(ValueError):
        client.transition_model_version_stage("me_flower_classifier", 1, stage="Bogus")


def test_small_mlflow_conveniences(tmp_path):
    """set_tags / log_text / log_artifacts / get_experiment_by_name / client.get_metric_history (MLflow spellings)."""
    import os

    tracking.set_tracking_uri(str(tmp_path / "mlruns"))
    exp = tracking.set_experiment("conv")
    assert tracking.get_experiment_by_name("conv").experiment_id == exp.experiment_id
    assert tracking.get_experiment_by_name("nope") is None
    d = tmp_path / "bundle"
    (d / "sub").mkdir(parents=True)
    (d / "a.txt").write_text("A")
    (d / "sub" / "b.txt").write_text("B")
    with tracking.start_run() as run:
        tracking.set_tags({"team": "vision", "stage": 2})
        tracking.log_text("hello", "notes/readme.txt")
        tracking.log_artifacts(str(d), "bundle")
        for s, v in enumerate((3.0, 2.0, 1.5)):
            tracking.log_metric("loss", v, step=s)
        rid = run.info.run_id
    r = tracking.get_run(rid)
    assert r.data.tags["team"] == "vision" and r.data.tags["stage"] == "2"
    root = r.info.artifact_uri
    assert open(os.path.join(root, "notes", "readme.txt")).read() == "hello"
    assert open(os.path.join(root, "bundle", "a.txt")).read() == "A"            # contents, not the directory itself
    assert open(os.path.join(root, "bundle", "sub", "b.txt")).read() == "B"
    hist = tracking.MlflowClient().get_metric_history(rid, "loss")
    assert [(h.step, h.value) for h in hist] == [(0, 3.0), (1, 2.0), (2, 1.5)] and all(h.key == "loss" for h in hist)
