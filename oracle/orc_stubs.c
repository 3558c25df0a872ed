/* temporary: games not yet restated */
#include "orc.h"
const orc_game_vt orc_vt_doudizhu = {0};
