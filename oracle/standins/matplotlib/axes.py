class Axes: pass
