from .base import Unit, Actor, Sensor
from .objects import Object, Box
from .robots import Robot, ArmRobot, LeggedRobot
from .sensors import CameraSensor
