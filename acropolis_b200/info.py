version = "0.1.0"
reference_version = "1.3.1"
description = "B200-native hot path of ACROPOLIS (photodisintegration of light elements)"
url = "https://acropolis.hepforge.org"
