"""Label loading (src/labels.rs:22-122): text, CSV (header heuristic), JSON (three shapes)."""
from __future__ import annotations

import csv
import io
import json
from typing import List

from .errors import LabelLoad, LabelParse
from .types import LabelFormat, ModelType


def load_labels_from_file(path: str, model_type: ModelType) -> List[str]:
    try:
        with open(path, "r", encoding="utf-8", newline="") as f:
            content = f.read()
    except (OSError, UnicodeDecodeError) as e:
        raise LabelLoad(str(path), str(e))
    return parse_labels(content, model_type.expected_label_format())


def parse_labels(content: str, fmt: LabelFormat) -> List[str]:
    if fmt is LabelFormat.Text:
        return parse_text_labels(content)
    if fmt is LabelFormat.Csv:
        return parse_csv_labels(content)
    return parse_json_labels(content)


def parse_text_labels(content: str) -> List[str]:           # labels.rs:42-48
    # str::lines splits on \n and strips one trailing \r; trim() then removes the rest
    return [t for t in (line.strip() for line in content.split("\n")) if t]


def looks_like_header(value: str) -> bool:                   # labels.rs:83-93
    lower = value.lower()
    return (lower in ("label", "species", "name", "class", "common_name", "scientific_name")
            or lower.startswith("inat") or lower.endswith("_fsd50k"))


def parse_csv_labels(content: str) -> List[str]:             # labels.rs:51-80
    labels: List[str] = []
    first_row = True
    try:
        reader = csv.reader(io.StringIO(content, newline=""), strict=False)
        for record in reader:
            if not record:          # csv crate skips empty lines
                continue
            label = record[0].strip()
            if first_row and looks_like_header(label):
                first_row = False
                continue
            first_row = False
            if label:
                labels.append(label)
    except csv.Error as e:
        raise LabelParse(str(e))
    return labels


def parse_json_labels(content: str) -> List[str]:            # labels.rs:96-122
    err = ("unrecognized JSON format: expected array of strings, {labels: [...]}, "
           "or [{name: ...}]")
    try:
        data = json.loads(content)
    except json.JSONDecodeError:
        raise LabelParse(err)
    if isinstance(data, list) and all(isinstance(x, str) for x in data):
        return list(data)
    if isinstance(data, dict) and isinstance(data.get("labels"), list) and \
            all(isinstance(x, str) for x in data["labels"]):
        return list(data["labels"])
    if isinstance(data, list) and all(isinstance(x, dict) for x in data):
        out = []
        for e in data:
            for key in ("name", "label", "species"):
                v = e.get(key)
                if isinstance(v, str):
                    out.append(v)
                    break
        if out:
            return out
    raise LabelParse(err)
