"""``model`` sections of the reference's shipped yaml configs (with input_dim/output_dim
injected the way wekws/bin/train.py:134-146 does).  Shapes per SURVEY.md section 8."""
import copy

_BASE = {
    # examples/hi_xiaowen/s0/conf/mdtc.yaml:28-38
    "mdtc": dict(hidden_dim=64, preprocessing=dict(type="linear"),
                 backbone=dict(type="mdtc", num_stack=4, stack_size=4, kernel_size=5, hidden_dim=64, causal=True)),
    # examples/hi_xiaowen/s0/conf/mdtc_small.yaml:28-38
    "mdtc_small": dict(hidden_dim=32, preprocessing=dict(type="linear"),
                       backbone=dict(type="mdtc", num_stack=3, stack_size=4, kernel_size=5, hidden_dim=32,
                                     causal=True)),
    # examples/hi_xiaowen/s0/conf/ds_tcn.yaml:27-36
    "ds_tcn": dict(hidden_dim=256, preprocessing=dict(type="linear"),
                   backbone=dict(type="tcn", ds=True, num_layers=4, kernel_size=8, dropout=0.1)),
    # examples/hi_xiaowen/s0/conf/tcn.yaml:26-35
    "tcn": dict(hidden_dim=64, preprocessing=dict(type="linear"),
                backbone=dict(type="tcn", ds=False, num_layers=4, kernel_size=8, dropout=0.1)),
    # examples/hi_xiaowen/s0/conf/gru.yaml:26-32
    "gru": dict(hidden_dim=128, preprocessing=dict(type="linear"), backbone=dict(type="gru", num_layers=2)),
    # examples/hi_xiaowen/s0/conf/fsmn_ctc.yaml:36-55 (input_dim 400 = 80-dim fbank x context 2+1+2, output_dim 2599)
    "fsmn": dict(hidden_dim=128, preprocessing=dict(type="none"),
                 backbone=dict(type="fsmn", input_affine_dim=140, num_layers=4, linear_dim=250, proj_dim=128,
                               left_order=10, right_order=2, left_stride=1, right_stride=1, output_affine_dim=140),
                 classifier=dict(type="identity", dropout=0.1), activation=dict(type="identity")),
}


def model_config(name: str, input_dim: int = 80, output_dim: int = 1, activation: str = "sigmoid",
                 cmvn_file=None, norm_var: bool = True) -> dict:
    cfg = copy.deepcopy(_BASE[name])
    cfg["input_dim"] = input_dim
    cfg["output_dim"] = output_dim
    if activation == "identity":          # examples/hi_xiaowen/s0/conf/ds_tcn_ctc.yaml:41-42
        cfg["activation"] = dict(type="identity")
    if cmvn_file is not None:
        cfg["cmvn"] = dict(cmvn_file=cmvn_file, norm_var=norm_var)
    return cfg


MODEL_NAMES = tuple(_BASE)
