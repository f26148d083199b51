"""TEST INFRASTRUCTURE — CPU oracle of the reference's PCM->spectrum->pixels path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (glava_b200) never does.
"""
