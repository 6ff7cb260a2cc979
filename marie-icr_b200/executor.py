"""TextExtractionExecutorB200 — the serving-side wrapper around the engine: mirror of
TextExtractionExecutor.extract (marie/executor/text/text_extraction_executor.py:125-260) and of the `ocr_frames` pipeline
step with its on-disk JSON cache (marie/pipe/components.py:569-656, marie/utils/json.py:19-40).

Not rebuilt: the Jina executor base class, DocList / asset storage, S3 / PostgreSQL persistence, the other pipeline
steps (segmentation, classification, overlays) — out of scope.  `extract` takes the frames (or a file
path) directly where the reference resolves `doc.asset_key` through its StorageManager.
"""
import json
import os

import numpy as np

from .ingest import ensure_max_page_size, frames_from_file, hash_frames_fast
from .plugin_api import CoordinateFormat, PSMode


class EnhancedJSONEncoder(json.JSONEncoder):
    """numpy scalars / arrays -> plain JSON (marie/numpyencoder.py:15-50, marie/utils/json.py EnhancedJSONEncoder)"""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.bool_):
            return bool(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def store_json_object(results, json_path):
    """marie/utils/json.py:19-30 — same formatting, so the files are byte-comparable with the reference's"""
    with open(os.path.expanduser(json_path), "w") as f:
        json.dump(results, f, sort_keys=False, separators=(",", ": "), ensure_ascii=True, indent=2, cls=EnhancedJSONEncoder)


def load_json_file(filename):
    with open(os.path.expanduser(filename), "r", encoding="utf-8") as f:
        return json.load(f)


def ocr_frames(ocr_engines, ref_id, frames, root_asset_dir, force=False, ps_mode=PSMode.SPARSE, coord_format=CoordinateFormat.XYWH,
               regions=None, runtime_conf=None, engine_name="default"):
    """marie/pipe/components.py:569-656: run the engine unless `<root_asset_dir>/results/<prefix>.json` (or
    `<prefix>.regions.json`) exists and `force` is off; the runtime configuration may pick another engine / force."""
    output_dir = os.path.join(root_asset_dir, "results")
    os.makedirs(output_dir, exist_ok=True)
    prefix = ref_id.split("/")[-1].split(".")[0]
    engine = ocr_engines[engine_name]
    if regions and len(regions) > 0:
        engine = ocr_engines["best"] if "best" in ocr_engines else engine
    if runtime_conf is not None:
        conf = runtime_conf.get("ocr", {})
        node = "document" if "document" in conf else ("region" if "region" in conf else "document")
        if node in conf:
            if "engine" in conf[node] and conf[node]["engine"] in ocr_engines:
                engine = ocr_engines[conf[node]["engine"]]
            if "force" in conf[node]:
                force = str(conf[node]["force"]).lower() in ("y", "yes", "t", "true", "on", "1")
    json_path = os.path.join(output_dir, f"{prefix}.regions.json" if regions and len(regions) > 0 else f"{prefix}.json")
    if force or not os.path.exists(json_path):
        results = engine.extract(frames, ps_mode, coord_format, regions)
        store_json_object(results, json_path)
    else:
        results = load_json_file(json_path)
    return results


def _payload_value(payload, key, default=None):
    """value_from_payload_or_args: the value may sit in the payload itself or in its 'args' object"""
    if key in payload:
        return payload[key]
    if "args" in payload and key in payload["args"]:
        return payload["args"][key]
    return default


class TextExtractionExecutorB200:
    """extract(frames | path, parameters) -> {"status", "runtime_info", "metadata"} with the reference's request
    validation, parameter handling and error envelope (:125-260)."""

    def __init__(self, engine=None, workspace="/tmp/marie", name="TextExtractionExecutorB200", **engine_kwargs):
        if engine is None:
            from .engine import OcrEngineB200
            engine = OcrEngineB200(**engine_kwargs)
        self.engines = {"default": engine, "best": engine}
        self.workspace = workspace
        self.show_error = True
        self.runtime_info = {"name": self.__class__.__name__, "instance_name": name, "model": "", "workspace": workspace,
                             "use_cuda": True}

    def extract(self, docs, parameters, *args, **kwargs):
        if docs is None or (hasattr(docs, "__len__") and len(docs) == 0):
            return {"error": "empty payload"}
        src_frames = frames_from_file(docs) if isinstance(docs, (str, os.PathLike)) else list(docs)
        changed, frames = ensure_max_page_size(src_frames)
        if parameters is None or "job_id" not in parameters:
            raise ValueError("Job ID is not present in parameters")
        queue_id = parameters.get("queue_id", "0000-0000-0000-0000")
        try:
            if "payload" not in parameters or parameters["payload"] is None:
                return {"error": "empty payload"}
            payload = parameters["payload"]
            regions = payload["regions"] if "regions" in payload else []
            for region in regions:                                   # :160-167
                region["id"] = f'{int(region["id"])}'
                for k in ("x", "y", "w", "h", "pageIndex"):
                    region[k] = int(region[k])
            coordinate_format = CoordinateFormat.from_value(_payload_value(payload, "format", default="xywh"))
            pms_mode = PSMode.from_value(_payload_value(payload, "mode", default=""))
            ref_id = parameters.get("ref_id") or hash_frames_fast(frames)
            ref_type = parameters.get("ref_type") or "extract"
            runtime_conf = {}
            for feature in payload.get("features", []):
                if feature.get("type") == "pipeline":
                    runtime_conf = feature
            include_ocr = _payload_value(payload, "return_ocr", default=False)
            root_asset_dir = os.path.join(self.workspace, "generators", str(ref_type), str(ref_id).split("/")[-1].split(".")[0])
            ocr = ocr_frames(self.engines, str(ref_id), frames, root_asset_dir, force=bool(parameters.get("force", False)),
                             ps_mode=pms_mode, coord_format=coordinate_format, regions=regions, runtime_conf=runtime_conf)
            metadata = {"ref_id": ref_id, "ref_type": ref_type, "job_id": parameters.get("job_id"), "queue_id": queue_id,
                        "pipeline": "default", "pages": f"{len(frames)}", "ocr": ocr}
            if not include_ocr:
                del metadata["ocr"]
            return json.loads(json.dumps({"status": "succeeded", "runtime_info": self.runtime_info, "metadata": metadata},
                                         cls=EnhancedJSONEncoder))
        except BaseException as error:                               # the reference's error envelope (:243-253)
            return {"status": "error", "runtime_info": self.runtime_info, "error": (str(error),) if self.show_error else "inference exception"}
