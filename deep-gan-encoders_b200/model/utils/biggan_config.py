"""BigGAN configuration object -- same attributes / constructors as the reference's `model/utils/biggan_config.py`
(:11-71); defaults are the 128x128 model, `layers` tuples are (up-sample?, in multiple, out multiple)."""
import copy
import json


class BigGANConfig(object):
    def __init__(self, output_dim=128, z_dim=128, class_embed_dim=128, channel_width=128, num_classes=1000,
                 layers=((False, 16, 16), (True, 16, 16), (False, 16, 16), (True, 16, 8), (False, 8, 8), (True, 8, 4),
                         (False, 4, 4), (True, 4, 2), (False, 2, 2), (True, 2, 1)),
                 attention_layer_position=8, eps=1e-4, n_stats=51):
        self.output_dim = output_dim
        self.z_dim = z_dim
        self.class_embed_dim = class_embed_dim
        self.channel_width = channel_width
        self.num_classes = num_classes
        self.layers = [tuple(l) for l in layers]
        self.attention_layer_position = attention_layer_position
        self.eps = eps
        self.n_stats = n_stats

    @classmethod
    def from_dict(cls, json_object):
        config = BigGANConfig()
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding='utf-8') as reader:
            return cls.from_dict(json.loads(reader.read()))

    def __repr__(self):
        return str(self.to_json_string())

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"
