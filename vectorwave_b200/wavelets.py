"""Wavelet definitions the MODWT path consumes: filter tables and BoundaryMode.

Mirrors CORE/api: `Wavelet.lowPassDecomposition()/highPassDecomposition()/lowPassReconstruction()/
highPassReconstruction()`, singletons `Daubechies.DB4`, `Symlet.SYM8`, `Coiflet.COIF5`, `Haar.INSTANCE`
(Haar.java:39-43, Daubechies.java:61-160, Symlet.java:91-96,156-161, Coiflet.java:63-178).
Orthogonal: reconstruction == decomposition filters; QMF g[i] = (-1)^i h[L-1-i]
(Daubechies.java:323-330).  The engine never hard-codes tables: filters cross the C ABI as data.
"""
import enum
import math

import numpy as np


class BoundaryMode(enum.Enum):
    """CORE/api/BoundaryMode.java:26-50.  Values are the native vw_boundary integers."""
    PERIODIC = 0
    ZERO_PADDING = 1
    SYMMETRIC = 2
    CONSTANT = 3


class Wavelet:
    """Orthogonal discrete wavelet (CORE/api/OrthogonalWavelet.java defaults)."""

    def __init__(self, name, lowpass, vanishing_moments=0):
        self._name = name
        self._h = np.array(lowpass, dtype=np.float64)
        self._vm = vanishing_moments

    def name(self):
        return self._name

    def lowPassDecomposition(self):
        return self._h.copy()

    def highPassDecomposition(self):
        h = self._h
        sign = np.where(np.arange(h.size) % 2 == 0, 1.0, -1.0)
        return sign * h[::-1]

    def lowPassReconstruction(self):
        return self.lowPassDecomposition()

    def highPassReconstruction(self):
        return self.highPassDecomposition()

    def vanishingMoments(self):
        return self._vm

    def __repr__(self):
        return f"Wavelet({self._name}, L={self._h.size})"


_S = 1.0 / math.sqrt(2)


class Haar(Wavelet):
    """CORE/api/Haar.java:39-43; `Haar()` instances compare by type like the Java record."""
    INSTANCE = None

    def __init__(self):
        super().__init__("Haar", [_S, _S], 1)

    def highPassDecomposition(self):
        return np.array([_S, -_S])


Haar.INSTANCE = Haar()


class Daubechies:
    DB2 = Wavelet("db2", [0.4829629131445341, 0.8365163037378079, 0.2241438680420134, -0.1294095225512603], 2)
    DB4 = Wavelet("db4", [0.2303778133088964, 0.7148465705529154, 0.6308807679298587, -0.0279837693982488,
                          -0.1870348117190931, 0.0308413818355607, 0.0328830116668852, -0.0105974017850690], 4)
    DB6 = Wavelet("db6", [0.1115407433501094, 0.4946238903984530, 0.7511339080210954, 0.3152503517091980,
                          -0.2262646939654399, -0.1297668675672624, 0.0975016055873224, 0.0275228655303053,
                          -0.0315820393174862, 0.0005538422011614, 0.0047772575109455, -0.0010773010853085], 6)
    DB8 = Wavelet("db8", [0.0544158422431049, 0.3128715909143031, 0.6756307362972904, 0.5853546836541907,
                          -0.0158291052563816, -0.2840155429615702, 0.0004724845739124, 0.1287474266204837,
                          -0.0173693010018083, -0.0440882539307952, 0.0139810279173995, 0.0087460940474061,
                          -0.0048703529934518, -0.0003917403733770, 0.0006754494064506, -0.0001174767841248], 8)
    DB10 = Wavelet("db10", [0.0266700579005546, 0.1881768000776347, 0.5272011889317202, 0.6884590394536250,
                            0.2811723436605715, -0.2498464243271598, -0.1959462743772862, 0.1273693403357932,
                            0.0930573646035547, -0.0713941471663501, -0.0294575368218399, 0.0332126740593612,
                            0.0036065535669870, -0.0107331754833007, 0.0013953517470688, 0.0019924052951925,
                            -0.0006858566949564, -0.0001164668551285, 0.0000935886703202, -0.0000132642028945], 10)


class Symlet:
    SYM4 = Wavelet("sym4", [0.03222310060407815, -0.01260396726226383, -0.09921954357695636, 0.29785779560553225,
                            0.80373875180591614, 0.49761866763256292, -0.02963552764596039, -0.07576571478935668], 4)
    SYM8 = Wavelet("sym8", [-0.003382415951359, -0.000542132331635, 0.031695087810979, 0.007607487324918,
                            -0.143294238350810, -0.061273359067938, 0.481359651258372, 0.777185751700574,
                            0.364441894835509, -0.051945838107658, -0.027219029168752, 0.049137179673713,
                            0.003808752013903, -0.014952258336792, -0.000302920514551, 0.001889950332768], 8)


class Coiflet:
    COIF2 = Wavelet("coif2", [-0.0007205494453645, -0.0018232088709132, 0.0056211431711065, 0.0235962077162017,
                              -0.0594274367855454, -0.0764421423447531, 0.4170051844216925, 0.8127236354455423,
                              0.3861100668250532, -0.0673725547219630, -0.0414649367817581, 0.0164064277978058], 4)
    COIF3 = Wavelet("coif3", [-0.0000345997728362, -0.0000709833031381, 0.0004662169601129, 0.0011175187708906,
                              -0.0025745176887502, -0.0090079761366615, 0.0158805448636158, 0.0345550275730615,
                              -0.0823019271068856, -0.0717998216193117, 0.4284834763776168, 0.7937772226256169,
                              0.4051769024096150, -0.0611233900026726, -0.0657719112818552, 0.0234526961418362,
                              0.0077825964273254, -0.0037935128644910], 6)
    COIF5 = Wavelet("coif5", [-0.0000000960401011, -0.0000001623799517, 0.0000020612203986, 0.0000037007277113,
                              -0.0000212702216725, -0.0000412198619243, 0.0001403563281237, 0.0003018579416682,
                              -0.0006375589261259, -0.0016616273039299, 0.0024315754425383, 0.0067615202206204,
                              -0.0091595073386762, -0.0197583916009655, 0.0326747994670574, 0.0412875304721178,
                              -0.1055631513073372, -0.0620377515749820, 0.4379823066591634, 0.7742936228603274,
                              0.4215712667307543, -0.0520466702535548, -0.0919215880600861, 0.0281697442705324,
                              0.0234083221189278, -0.0101315848469003, -0.0041593126275786, 0.0021782943778457,
                              0.0003585777411618, -0.0002120818620675], 10)


REGISTRY = {
    "haar": Haar.INSTANCE, "db2": Daubechies.DB2, "db4": Daubechies.DB4, "db6": Daubechies.DB6,
    "db8": Daubechies.DB8, "db10": Daubechies.DB10, "sym4": Symlet.SYM4, "sym8": Symlet.SYM8,
    "coif2": Coiflet.COIF2, "coif3": Coiflet.COIF3, "coif5": Coiflet.COIF5,
}


def get_wavelet(name):
    """CORE/api/WaveletRegistry.java:25-50 -- hands out the shared static instances."""
    return REGISTRY[name.lower()]
