"""Scene-JSON access (mirror of eng/configer_builder.py:3-38 of the reference; same class and method names)."""
import json


class SimConfiger:
    _SECTIONS = {"get_materials": "Materials", "get_blocks": "Blocks", "get_bodies": "Bodies", "get_motions": "Motions"}

    def __init__(self, scene_file_path=None, config=None) -> None:
        """``scene_file_path``: path of a tiSPHi scene JSON.  ``config``: an already loaded dict (extension)."""
        if config is not None:
            self.config = config
        else:
            with open(scene_file_path, "r") as f:
                self.config = json.load(f)
        print("\n========== CONFIGURE LOADED ==========")

    def get_cfg(self, name, enforce_exist=False):
        # a missing key is a KeyError, exactly like the reference (configer_builder.py:11-14)
        section = self.config["Configuration"]
        if enforce_exist:
            assert name in section
        return section[name]

    def get_opt(self, name, default=None):
        """Optional keys added by this engine (``precision``, ``wcFresh``, ...) - never required."""
        return self.config["Configuration"].get(name, default)

    def _section(self, key):
        return self.config[key] if key in self.config else []

    def get_materials(self):
        return self._section("Materials")

    def get_blocks(self):
        return self._section("Blocks")

    def get_bodies(self):
        return self._section("Bodies")

    def get_motions(self):
        return self._section("Motions")
