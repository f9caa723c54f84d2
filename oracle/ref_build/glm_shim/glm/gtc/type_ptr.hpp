// empty: the reference includes this header but uses nothing from it (ewa_project.cu:9)
#pragma once
