// Link seam for the reference's HM builds: same name as
// hevc/hm_common/c++/source_common/interface_c_python.h (included from TComPrediction.h:46), first on the
// include path.  The reference embeds CPython 2.7 only to unpickle the training mean
// (TComPrediction.cpp:180-236); this header declares the few CPython names used there, implemented in
// pnn_hm_shim.cpp by a reader of the pickled float (protocol 0/1/2) or of a plain text float.
#ifndef INTERFACE_C_PYTHON_H
#define INTERFACE_C_PYTHON_H

#include <assert.h>
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

struct PyObject {
    bool is_float;
    double value;
};

void Py_Initialize();
int Py_IsInitialized();
void Py_Finalize();
void Py_DECREF(PyObject* object);
int PyFloat_CheckExact(PyObject* object);
double PyFloat_AsDouble(PyObject* object);
PyObject* PyErr_Occurred();
void PyErr_Print();

// reference interface_c_python.h:21-50
int append_sys_path(const std::string& path_to_additional_directory);
PyObject* get_callable(const std::string& name_file, const std::string& name_function);
PyObject* load_via_pickle(PyObject* python_function, const std::string& path_to_file);

#endif
