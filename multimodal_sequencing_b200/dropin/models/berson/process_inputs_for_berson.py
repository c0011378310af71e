"""models/berson/process_inputs_for_berson.py of the reference (13-79, 246-261): same function names and
return dict; the expansion itself is the vectorised host logic of multimodal_sequencing_b200.engine."""
from multimodal_sequencing_b200.engine import pairs_generator, prepare_pairs  # noqa: F401


def prepare_berson_inputs(inputs, tokenizer, args=None):
    cls_id = tokenizer.convert_tokens_to_ids(tokenizer.cls_token)
    sep_id = tokenizer.convert_tokens_to_ids(tokenizer.sep_token)
    pad_id = tokenizer.convert_tokens_to_ids(tokenizer.pad_token)
    images = inputs["images"] if "images" in inputs and inputs["images"] is not None else None
    pb = prepare_pairs(inputs["input_ids"], inputs["labels"], args.max_story_length, images, cls_id, sep_id, pad_id)
    d = pb.reference_dict(materialize_images=True)
    d["_pair_batch"] = pb  # compact form (unique images + index table) for the fast path
    for k, v in d.items():
        if hasattr(v, "to") and k not in ("cuda", "_pair_batch"):
            d[k] = v.to(args.device)
    return d
