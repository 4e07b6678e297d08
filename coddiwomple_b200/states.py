"""States module — abstract interfaces with the reference's names (coddiwomple/states.py:11-53)."""


class ParticleState():
    """Generalized ParticleState object to hold a particle state x and potentially other attributes"""
    pass


class PDFState():
    """Generalized PDFState object that defines a parametrizable probability distribution function"""

    def reduced_potential(self, particle_state):
        """compute the 'reduced potential energy' of a ParticleState"""
        pass

    def set_parameters(self, parameters):
        """update the pdf_state parameters in place"""
        pass

    def get_parameters(self):
        """return the current parameters as {name: value}"""
        pass
