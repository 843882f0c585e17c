"""The one piece of Envs/ai2thor/RL_env_VAR.py the triplet path needs: `Task`
(Envs/ai2thor/RL_env_VAR.py:23-35), the (location, object, action) key that VARDataset's task list
(dataset.py:17-29) and audioLoader.getAudioFromTask (Envs/audioLoader.py:223-237) pass around.
The Unity simulator environment itself (`RLEnvVAR`) is a caller of the path and stays with the
reference (out of scope)."""


class Task(object):
    def __init__(self, loc, obj, act):
        self.loc, self.obj, self.act = loc, obj, act

    def _key(self):
        return (self.loc, self.obj, self.act)

    def __eq__(self, other):
        return self._key() == (other.loc, other.obj, other.act)

    def __ne__(self, other):
        return not (self == other)

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return "Task(loc=%r, obj=%r, act=%r)" % self._key()
