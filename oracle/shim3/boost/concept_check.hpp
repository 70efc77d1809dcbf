// empty stand-in (boost/concept_check.hpp is included by the reference but nothing of it is used)
