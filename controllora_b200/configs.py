"""ControlLoRA configurations shipped by the reference as configs/*.json, expressed as overrides of the constructor
defaults (models.py:620-667).  `ControlLoRA.from_config(NAMED["diffusiondb-canny-v2"])` equals loading the reference's
configs/diffusiondb-canny-v2.json; the JSON files themselves load unchanged through `ControlLoRA.from_config(path)`."""

_V2 = dict(lora_control_version=2, lora_pre_conv_skipped=True, lora_concat_hidden=True, lora_control_channels=[256, 256, 256],
           lora_control_self_add=False, lora_key_states_skipped=True, lora_value_states_skipped=True,
           lora_output_states_skipped=False)

NAMED = {
    "base": {},
    "fill50k": {},
    "diffusiondb-canny": {},
    "mpii-pose": {},
    "diffusiondb-canny-v2": dict(_V2),
    "mpii-pose-v2": dict(_V2),
    "post-add": dict(lora_post_add=True),
    "danbooru-sketch": dict(lora_control_channels=[256, 256, 256], lora_control_rank=256, lora_control_self_add=False,
                            lora_concat_hidden=True, lora_pre_conv_skipped=True),
}


def wire_processors(unet, control_lora):
    """The reference's processor wiring (train_text_to_image_control_lora.py:469-487): walk `unet.attn_processors` in key
    order and pop processors from `control_lora.lora_layers[control_id]`."""
    n_ch = len(unet.config.block_out_channels)
    ids = list(range(n_ch))
    pools = [list(l) for l in control_lora.lora_layers]
    procs = {}
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            cid = ids[-1]
        elif name.startswith("up_blocks"):
            cid = list(reversed(ids))[int(name[len("up_blocks."):].split(".")[0])]
        else:
            cid = ids[int(name[len("down_blocks."):].split(".")[0])]
        if pools[cid]:
            procs[name] = pools[cid].pop(0)
    unet.set_attn_processor(procs)
    return procs
