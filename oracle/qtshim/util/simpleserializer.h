// Oracle shim: sdrbase/dsp/filerecord.cpp includes util/simpleserializer.h without using it; the real header needs QMap and
// QByteArray.  This empty stand-in (found first through -Iqtshim) lets the UNMODIFIED filerecord.cpp compile.
#ifndef ORACLE_QTSHIM_SIMPLESERIALIZER_H
#define ORACLE_QTSHIM_SIMPLESERIALIZER_H
#endif
